// fastnoise2_b200 -- intrusive reference-counted node handle.
//
// Public surface follows the reference's FastNoise::SmartNode<T> (Utility/SmartNode.h:70-316):
// copy/move between compatible node types, DynamicCast, get/use_count/unique/swap, comparison and
// std::hash. The count lives in the node (Generator::mRefs); the last handle destroys the node and,
// with it, any device program that was lowered from it. The reference's bump-pool allocator
// (src/FastNoise/SmartNode.cpp) is not reproduced -- node allocation is not on the hot path here.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <type_traits>
#include <utility>

namespace FastNoise
{
    class FASTNOISE_API SmartNodeManager
    {
    public:
        static constexpr std::uint64_t kInvalidReferenceId = 0;
        SmartNodeManager() = delete;
        /// Kept for source compatibility (reference SmartNode.h:31); nodes are individually allocated here.
        static void SetMemoryPoolSize( std::uint32_t size );
    };

    namespace Internal
    {
        FASTNOISE_API std::int32_t NodeRefAdd( const Generator* node, std::int32_t delta ); // returns the previous count
        FASTNOISE_API void NodeDestroy( const Generator* node );
    }

    template<typename T>
    class SmartNode
    {
        static_assert( std::is_base_of<Generator, T>::value, "SmartNode should only be used for FastNoise node classes" );

        T* mNode = nullptr;

        static void Acquire( const Generator* node ) noexcept
        {
            if( node ) Internal::NodeRefAdd( node, 1 );
        }

        void Drop() noexcept
        {
            if( mNode && Internal::NodeRefAdd( mNode, -1 ) == 1 )
            {
                Internal::NodeDestroy( mNode );
            }
            mNode = nullptr;
        }

        template<typename U>
        friend class SmartNode;

    public:
        /// Adopt a raw node pointer (adds a reference). Used by New<T>() and the C ABI.
        static SmartNode FromRaw( T* node ) noexcept
        {
            SmartNode out;
            out.mNode = node;
            Acquire( node );
            return out;
        }

        template<typename U>
        static SmartNode DynamicCast( const SmartNode<U>& node )
        {
            return FromRaw( dynamic_cast<T*>( node.mNode ) );
        }

        constexpr SmartNode( std::nullptr_t = nullptr ) noexcept {}

        SmartNode( const SmartNode& other ) noexcept : mNode( other.mNode ) { Acquire( mNode ); }

        template<typename U, typename = typename std::enable_if<std::is_convertible<U*, T*>::value>::type>
        SmartNode( const SmartNode<U>& other ) noexcept : mNode( other.mNode ) { Acquire( mNode ); }

        /// aliasing form: share ownership with `owner` but point at `node` (already cast by the caller)
        template<typename U>
        SmartNode( const SmartNode<U>&, T* node ) noexcept : mNode( node ) { Acquire( mNode ); }

        SmartNode( SmartNode&& other ) noexcept : mNode( other.mNode ) { other.mNode = nullptr; }

        template<typename U, typename = typename std::enable_if<std::is_convertible<U*, T*>::value>::type>
        SmartNode( SmartNode<U>&& other ) noexcept : mNode( other.mNode ) { other.mNode = nullptr; }

        ~SmartNode() { Drop(); }

        SmartNode& operator=( const SmartNode& other ) noexcept
        {
            SmartNode copy( other );
            swap( copy );
            return *this;
        }

        template<typename U, typename = typename std::enable_if<std::is_convertible<U*, T*>::value>::type>
        SmartNode& operator=( const SmartNode<U>& other ) noexcept
        {
            SmartNode copy( other );
            swap( copy );
            return *this;
        }

        SmartNode& operator=( SmartNode&& other ) noexcept
        {
            swap( other );
            return *this;
        }

        template<typename U, typename = typename std::enable_if<std::is_convertible<U*, T*>::value>::type>
        SmartNode& operator=( SmartNode<U>&& other ) noexcept
        {
            SmartNode moved( std::move( other ) );
            swap( moved );
            return *this;
        }

        template<typename U>
        friend bool operator==( const SmartNode& a, const SmartNode<U>& b ) noexcept { return a.get() == b.get(); }
        template<typename U>
        friend bool operator!=( const SmartNode& a, const SmartNode<U>& b ) noexcept { return a.get() != b.get(); }

        T& operator*() const noexcept { return *mNode; }
        T* operator->() const noexcept { return mNode; }
        explicit operator bool() const noexcept { return mNode != nullptr; }
        T* get() const noexcept { return mNode; }

        void reset() noexcept { Drop(); }
        void swap( SmartNode& other ) noexcept { std::swap( mNode, other.mNode ); }

        long use_count() const noexcept { return mNode ? (long)Internal::NodeRefAdd( mNode, 0 ) : 0; }
        bool unique() const noexcept { return use_count() == 1; }
    };
}

namespace std
{
    template<typename T>
    struct hash<FastNoise::SmartNode<T>>
    {
        size_t operator()( const FastNoise::SmartNode<T>& node ) const noexcept
        {
            return std::hash<const void*>()( node.get() );
        }
    };
}
