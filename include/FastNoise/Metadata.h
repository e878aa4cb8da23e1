// fastnoise2_b200 -- node reflection and the encoded-node-tree format.
//
// Same public surface as the reference's FastNoise::Metadata / FastNoise::NodeData
// (include/FastNoise/Metadata.h:45-347): one Metadata record per node type, indexed by the node id
// that the encoded string carries (registration order, src/FastNoise/FastSIMD_Build.inl:86-143),
// with the member tables (variables / node lookups / hybrids) whose indices appear in the byte
// stream (src/FastNoise/Metadata.cpp:88-106). Here all 47 records are generated from one table
// (fastnoise2_b200/csrc/node_table.cpp) instead of per-class template constructors.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "Utility/Config.h"

namespace FastNoise
{
    struct NodeData;

    struct FASTNOISE_API Metadata
    {
        /// Read-only array view (the reference exposes the same begin/end/size/[] subset).
        template<typename T>
        class Vector
        {
        public:
            using const_iterator = const T*;
            const_iterator begin() const { return mBegin; }
            const_iterator end() const { return mBegin + mCount; }
            std::size_t size() const { return mCount; }
            const T& operator[]( std::size_t i ) const { return mBegin[i]; }

            Vector() = default;
            Vector( const T* first, std::size_t count ) : mBegin( first ), mCount( count ) {}

        private:
            const T* mBegin = nullptr;
            std::size_t mCount = 0;
        };

        struct Member
        {
            const char* name = "";
            const char* description = "";
            int dimensionIdx = -1; ///< 0..3 for per-dimension members, else -1
        };

        struct MemberVariable : Member
        {
            enum eType { EFloat, EInt, EEnum };

            union ValueUnion
            {
                float f;
                int i;
                ValueUnion( float v = 0 ) { f = v; }
                ValueUnion( int v ) { i = v; }
                operator float() const { return f; }
                operator int() const { return i; }
                bool operator==( const ValueUnion& rhs ) const { return i == rhs.i; }
            };

            eType type = EFloat;
            ValueUnion valueDefault, valueMin, valueMax;
            float valueUiDragSpeed = 0;
            Vector<const char*> enumNames;
            std::function<bool( Generator*, ValueUnion )> setFunc; ///< false if `Generator` is not this node type
        };

        struct MemberNodeLookup : Member
        {
            std::function<bool( Generator*, SmartNodeArg<> )> setFunc;
        };

        struct MemberHybrid : Member
        {
            float valueDefault = 0;
            float valueUiDragSpeed = 0;
            std::function<bool( Generator*, float )> setValueFunc;
            std::function<bool( Generator*, SmartNodeArg<> )> setNodeFunc;
        };

        using node_id = std::uint8_t;

        static const Vector<const Metadata*>& GetAll();
        static const Metadata* GetFromId( node_id nodeId );

        template<typename T>
        static const Metadata& Get()
        {
            static_assert( std::is_base_of<Generator, T>::value, "Metadata::Get<T>: T must be a FastNoise node class" );
            return *GetFromId( T::kNodeId );
        }

        /// Encode a NodeData tree (reference: Metadata.cpp:161-308). Empty string on failure.
        static std::string SerialiseNodeData( NodeData* nodeData, bool fixUp = false );
        /// Decode into NodeData records owned by `nodeDataOut`; returns the root or nullptr (Metadata.cpp:508-632).
        static NodeData* DeserialiseNodeData( const char* serialisedBase64NodeData, std::vector<std::unique_ptr<NodeData>>& nodeDataOut );

        static std::string FormatMetadataNodeName( const Metadata* metadata, bool removeGroups = false );
        static std::string FormatMetadataMemberName( const Member& member );

        /// Instantiate a node of this type with default parameters.
        SmartNode<> CreateNode( FastSIMD::FeatureSet maxFeatureSet = FastSIMD::FeatureSet::Max ) const;

        node_id id = 0;
        Vector<MemberVariable> memberVariables;
        Vector<MemberNodeLookup> memberNodeLookups;
        Vector<MemberHybrid> memberHybrids;
        Vector<const char*> groups;
        const char* name = "";
        const char* description = "";
        const char* formattedName = nullptr;
    };

    /// Editable description of one node instance; what the serialiser consumes and the NodeData decoder produces.
    struct FASTNOISE_API NodeData
    {
        explicit NodeData( const Metadata* data );

        const Metadata* metadata;
        std::vector<Metadata::MemberVariable::ValueUnion> variables;
        std::vector<NodeData*> nodeLookups;
        std::vector<std::pair<NodeData*, float>> hybrids;

        bool operator==( const NodeData& rhs ) const
        {
            return metadata == rhs.metadata && variables == rhs.variables && nodeLookups == rhs.nodeLookups && hybrids == rhs.hybrids;
        }
    };
}
