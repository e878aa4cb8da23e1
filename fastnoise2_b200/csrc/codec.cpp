// fastnoise2_b200 -- the encoded-node-tree text format.
//
// Format (restated from the reference, src/FastNoise/Metadata.cpp:88-158,161-308,310-506 and
// src/FastNoise/Base64.h:12-192):
//   text   = base64 (RFC 4648 alphabet, '=' padding) where a run of 3..66 'A' may be written as '@'
//            followed by alphabet[run - 3]
//   bytes  = node
//   node   = 0xFF ref:u16le                      back-reference to the ref-th COMPLETED node (post-order)
//          | id:u8 member* end
//   member = tag(0,idx) value:4 bytes LE          variable idx = raw int/float bits
//          | tag(1,count) node{count}             all required node inputs, in table order
//          | tag(2,idx) node                      hybrid idx driven by a node
//          | tag(3,idx) value:4 bytes LE          hybrid idx constant
//   end    = tag(4,extra)                         closes this node and `extra` enclosing nodes
//   tag(t,i) = one byte, t in the low 3 bits and i in the high 5
// Variables and hybrid constants are only written when they differ from the table default.
#include <FastNoise/FastNoise.h>
#include <FastNoise/Metadata.h>

#include <cctype>
#include <cstring>
#include <unordered_map>
#include <unordered_set>

#include "internal.h"

using namespace FastNoise;

namespace
{
    const char kAlphabet[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";

    int SextetOf( unsigned char c )
    {
        if( c >= 'A' && c <= 'Z' ) return c - 'A';
        if( c >= 'a' && c <= 'z' ) return c - 'a' + 26;
        if( c >= '0' && c <= '9' ) return c - '0' + 52;
        if( c == '+' ) return 62;
        if( c == '/' ) return 63;
        return -1;
    }

    enum Tag : std::uint8_t { TagVariable = 0, TagLookups = 1, TagHybridNode = 2, TagHybridValue = 3, TagEnd = 4 };

    // ------------------------------------------------------------------ reading
    struct Reader
    {
        const std::vector<std::uint8_t>& bytes;
        std::size_t pos = 0;
        int pendingEnds = 0; // node-ends folded into the last end byte and not yet consumed

        bool Raw( void* out, std::size_t n )
        {
            if( pendingEnds > 0 || pos + n > bytes.size() ) return false;
            std::memcpy( out, bytes.data() + pos, n );
            pos += n;
            return true;
        }

        bool NextTag( int& type, int& index )
        {
            if( pendingEnds > 0 )
            {
                pendingEnds--;
                type = TagEnd;
                index = 0;
                return true;
            }
            std::uint8_t b;
            if( !Raw( &b, 1 ) ) return false;
            type = b & 7;
            index = b >> 3;
            if( type == TagEnd ) pendingEnds = index;
            return true;
        }
    };

    // One walker for both decode targets: live nodes and NodeData records.
    template<typename SINK>
    typename SINK::Handle ReadNode( Reader& in, SINK& sink )
    {
        using Handle = typename SINK::Handle;

        std::uint8_t id;
        if( !in.Raw( &id, 1 ) ) return Handle();

        if( id == 0xFF )
        {
            std::uint16_t ref;
            if( !in.Raw( &ref, 2 ) ) return Handle();
            return sink.Completed( ref );
        }

        const Metadata* meta = Metadata::GetFromId( id );
        if( !meta ) return Handle();

        Handle node = sink.Create( meta );
        if( !node ) return Handle();

        int type, index;
        if( !in.NextTag( type, index ) ) return Handle();

        while( type == TagVariable )
        {
            std::int32_t bits;
            if( !in.Raw( &bits, 4 ) ) return Handle();
            if( index < (int)meta->memberVariables.size() ) sink.Variable( node, meta, index, bits );
            if( !in.NextTag( type, index ) ) return Handle();
        }

        if( type == TagLookups )
        {
            const int declared = (int)meta->memberNodeLookups.size();
            for( int k = 0; k < index; k++ )
            {
                Handle child = ReadNode( in, sink ); // surplus inputs are still consumed
                if( k < declared )
                {
                    if( !sink.Lookup( node, meta, k, child ) ) return Handle();
                }
                else if( !child && SINK::kStrict )
                {
                    return Handle();
                }
            }
            for( int k = index; k < declared; k++ )
            {
                if( !sink.MissingLookup( node, meta, k ) ) return Handle();
            }
            if( !in.NextTag( type, index ) ) return Handle();
        }

        while( type != TagEnd )
        {
            if( type == TagHybridNode )
            {
                Handle child = ReadNode( in, sink );
                if( index < (int)meta->memberHybrids.size() && !sink.HybridNode( node, meta, index, child ) ) return Handle();
            }
            else if( type == TagHybridValue )
            {
                float value;
                if( !in.Raw( &value, 4 ) ) return Handle();
                if( index < (int)meta->memberHybrids.size() ) sink.HybridValue( node, meta, index, value );
            }
            if( !in.NextTag( type, index ) ) return Handle();
        }

        sink.Finish( node );
        return node;
    }

    struct LiveSink
    {
        using Handle = SmartNode<>;
        static constexpr bool kStrict = true;
        std::vector<SmartNode<>> done;

        Handle Completed( std::uint16_t ref ) { return ref < done.size() ? done[ref] : nullptr; }
        Handle Create( const Metadata* meta ) { return meta->CreateNode(); }
        void Variable( Handle& n, const Metadata*, int idx, std::int32_t bits ) { Impl::SetVariable( n.get(), idx, bits ); }
        bool Lookup( Handle& n, const Metadata*, int idx, Handle& child ) { return child && Impl::SetLookup( n.get(), idx, child.get() ); }
        bool MissingLookup( Handle& n, const Metadata*, int idx )
        {
            // a tree written by an older table: fill the new input with a Constant (Metadata.cpp:446-452)
            SmartNode<> filler = FastNoise::New<Constant>();
            return Impl::SetLookup( n.get(), idx, filler.get() );
        }
        bool HybridNode( Handle& n, const Metadata*, int idx, Handle& child ) { return child && Impl::SetHybridNode( n.get(), idx, child.get() ); }
        void HybridValue( Handle& n, const Metadata*, int idx, float v ) { Impl::SetHybridValue( n.get(), idx, v ); }
        void Finish( Handle& n ) { done.push_back( n ); }
    };

    struct DataSink
    {
        using Handle = NodeData*;
        static constexpr bool kStrict = false;
        std::vector<std::unique_ptr<NodeData>>& owned;
        std::vector<std::unique_ptr<NodeData>> building; // records whose end has not been read yet

        Handle Completed( std::uint16_t ref ) { return ref < owned.size() ? owned[ref].get() : nullptr; }
        Handle Create( const Metadata* meta )
        {
            building.emplace_back( new NodeData( meta ) );
            return building.back().get();
        }
        void Variable( Handle n, const Metadata*, int idx, std::int32_t bits ) { n->variables[idx] = Metadata::MemberVariable::ValueUnion( (int)bits ); }
        bool Lookup( Handle n, const Metadata*, int idx, Handle child ) { n->nodeLookups[idx] = child; return true; }
        bool MissingLookup( Handle, const Metadata*, int ) { return true; }
        bool HybridNode( Handle n, const Metadata*, int idx, Handle child ) { n->hybrids[idx].first = child; return true; }
        void HybridValue( Handle n, const Metadata*, int idx, float v ) { n->hybrids[idx].second = v; }
        void Finish( Handle n )
        {
            for( auto it = building.begin(); it != building.end(); ++it )
            {
                if( it->get() == n )
                {
                    owned.push_back( std::move( *it ) );
                    building.erase( it );
                    return;
                }
            }
        }
    };

    // ------------------------------------------------------------------ writing
    struct Writer
    {
        std::vector<std::uint8_t> bytes;
        int pendingEnds = 0;

        void FlushEnds()
        {
            if( pendingEnds > 0 )
            {
                bytes.push_back( (std::uint8_t)( TagEnd | ( ( pendingEnds - 1 ) << 3 ) ) );
                pendingEnds = 0;
            }
        }
        void Byte( std::uint8_t b ) { FlushEnds(); bytes.push_back( b ); }
        void Tag( int type, int index ) { Byte( (std::uint8_t)( type | ( index << 3 ) ) ); }
        void Word16( std::uint16_t v ) { Byte( v & 0xff ); Byte( v >> 8 ); }
        void Word32( std::uint32_t v ) { for( int k = 0; k < 4; k++ ) Byte( ( v >> ( 8 * k ) ) & 0xff ); }
        void End()
        {
            if( ++pendingEnds >= 32 ) FlushEnds(); // 5-bit counter
        }
    };

    bool InputAccepted( const Metadata* owner, bool hybrid, int index, const Metadata* candidate )
    {
        SmartNode<> probe = owner->CreateNode();
        SmartNode<> source = candidate->CreateNode();
        return hybrid ? owner->memberHybrids[index].setNodeFunc( probe.get(), source )
                      : owner->memberNodeLookups[index].setFunc( probe.get(), source );
    }

    // `ancestors` is passed by value so siblings do not see each other's descendants
    bool WriteNode( NodeData* data, bool fixUp, Writer& out, std::unordered_map<const NodeData*, std::uint16_t>& refs, std::unordered_set<const NodeData*> ancestors )
    {
        const Metadata* meta = data->metadata;
        if( !meta || data->variables.size() != meta->memberVariables.size() ||
            data->nodeLookups.size() != meta->memberNodeLookups.size() || data->hybrids.size() != meta->memberHybrids.size() )
        {
            return false;
        }

        if( fixUp ) // break cycles by disconnecting any input that is one of our ancestors
        {
            ancestors.insert( data );
            for( NodeData*& child : data->nodeLookups ) if( ancestors.count( child ) ) child = nullptr;
            for( auto& hybrid : data->hybrids ) if( ancestors.count( hybrid.first ) ) hybrid.first = nullptr;
        }

        auto known = refs.find( data );
        if( known != refs.end() )
        {
            out.Byte( 0xFF );
            out.Word16( known->second );
            return true;
        }

        out.Byte( meta->id );

        for( std::size_t k = 0; k < meta->memberVariables.size(); k++ )
        {
            if( data->variables[k].i != meta->memberVariables[k].valueDefault.i )
            {
                out.Tag( TagVariable, (int)k );
                out.Word32( (std::uint32_t)data->variables[k].i );
            }
        }

        if( meta->memberNodeLookups.size() )
        {
            out.Tag( TagLookups, (int)meta->memberNodeLookups.size() );
        }
        for( std::size_t k = 0; k < meta->memberNodeLookups.size(); k++ )
        {
            NodeData*& child = data->nodeLookups[k];
            if( fixUp && child && !InputAccepted( meta, false, (int)k, child->metadata ) )
            {
                child = nullptr;
                return false;
            }
            if( !child || !WriteNode( child, fixUp, out, refs, ancestors ) ) return false;
        }

        for( std::size_t k = 0; k < meta->memberHybrids.size(); k++ )
        {
            auto& hybrid = data->hybrids[k];
            if( !hybrid.first )
            {
                if( hybrid.second != meta->memberHybrids[k].valueDefault )
                {
                    out.Tag( TagHybridValue, (int)k );
                    out.Word32( (std::uint32_t)Impl::Bits( hybrid.second ) );
                }
                continue;
            }
            if( fixUp && !InputAccepted( meta, true, (int)k, hybrid.first->metadata ) )
            {
                hybrid.first = nullptr;
                return false;
            }
            out.Tag( TagHybridNode, (int)k );
            if( !WriteNode( hybrid.first, fixUp, out, refs, ancestors ) ) return false;
        }

        out.End();
        refs.emplace( data, (std::uint16_t)refs.size() );
        return true;
    }
}

namespace FastNoise
{
    namespace Impl
    {
        std::vector<std::uint8_t> Base64Decode( const char* text )
        {
            if( !text ) return {};

            // expand '@' runs into plain sextets first
            std::vector<std::uint8_t> sextets;
            std::size_t padding = 0;
            for( std::size_t k = 0; text[k]; k++ )
            {
                unsigned char c = (unsigned char)text[k];
                if( c == '@' )
                {
                    int extra = SextetOf( (unsigned char)text[++k] );
                    if( extra < 0 ) return {};
                    sextets.insert( sextets.end(), (std::size_t)extra + 3, 0 );
                }
                else if( c == '=' )
                {
                    sextets.push_back( 0 );
                    padding++;
                }
                else
                {
                    int v = SextetOf( c );
                    // the reference maps unknown characters to 64 and lets the shifts smear them; reject instead
                    if( v < 0 ) return {};
                    if( padding ) return {};
                    sextets.push_back( (std::uint8_t)v );
                }
            }
            if( sextets.empty() || sextets.size() % 4 != 0 || padding > 2 ) return {};

            std::vector<std::uint8_t> bytes;
            bytes.reserve( sextets.size() / 4 * 3 );
            for( std::size_t k = 0; k < sextets.size(); k += 4 )
            {
                std::uint32_t triple = ( sextets[k] << 18 ) | ( sextets[k + 1] << 12 ) | ( sextets[k + 2] << 6 ) | sextets[k + 3];
                bytes.push_back( ( triple >> 16 ) & 0xff );
                bytes.push_back( ( triple >> 8 ) & 0xff );
                bytes.push_back( triple & 0xff );
            }
            bytes.resize( bytes.size() - padding );
            return bytes;
        }

        std::string Base64Encode( const std::vector<std::uint8_t>& bytes )
        {
            // plain base64 first, run-length pass second
            std::string plain;
            std::size_t k = 0;
            for( ; k + 2 < bytes.size(); k += 3 )
            {
                std::uint32_t triple = ( bytes[k] << 16 ) | ( bytes[k + 1] << 8 ) | bytes[k + 2];
                plain += kAlphabet[( triple >> 18 ) & 63];
                plain += kAlphabet[( triple >> 12 ) & 63];
                plain += kAlphabet[( triple >> 6 ) & 63];
                plain += kAlphabet[triple & 63];
            }
            if( k < bytes.size() )
            {
                std::uint32_t triple = bytes[k] << 16;
                bool two = k + 1 < bytes.size();
                if( two ) triple |= bytes[k + 1] << 8;
                plain += kAlphabet[( triple >> 18 ) & 63];
                plain += kAlphabet[( triple >> 12 ) & 63];
                plain += two ? kAlphabet[( triple >> 6 ) & 63] : '=';
                plain += '=';
            }

            std::string packed;
            for( std::size_t p = 0; p < plain.size(); )
            {
                std::size_t run = 0;
                while( p + run < plain.size() && plain[p + run] == 'A' ) run++;
                if( run >= 3 )
                {
                    // longest run one '@' pair can carry is 63 + 3
                    std::size_t take = run > 66 ? 66 : run;
                    packed += '@';
                    packed += kAlphabet[take - 3];
                    p += take;
                }
                else if( run > 0 )
                {
                    packed.append( run, 'A' );
                    p += run;
                }
                else
                {
                    packed += plain[p++];
                }
            }
            return packed;
        }
    }

    SmartNode<> NewFromEncodedNodeTree( const char* encodedNodeTreeString, FastSIMD::FeatureSet maxFeatureSet )
    {
        if( maxFeatureSet == FastSIMD::FeatureSet::Invalid ) return nullptr;
        std::vector<std::uint8_t> bytes = Impl::Base64Decode( encodedNodeTreeString );
        Reader reader { bytes };
        LiveSink sink;
        return ReadNode( reader, sink );
    }

    NodeData* Metadata::DeserialiseNodeData( const char* serialisedBase64NodeData, std::vector<std::unique_ptr<NodeData>>& nodeDataOut )
    {
        std::vector<std::uint8_t> bytes = Impl::Base64Decode( serialisedBase64NodeData );
        Reader reader { bytes };
        DataSink sink { nodeDataOut, {} };
        return ReadNode( reader, sink );
    }

    std::string Metadata::SerialiseNodeData( NodeData* nodeData, bool fixUp )
    {
        if( !nodeData ) return "";
        Writer out;
        std::unordered_map<const NodeData*, std::uint16_t> refs;
        if( !WriteNode( nodeData, fixUp, out, refs, {} ) ) return "";
        out.FlushEnds();
        return Impl::Base64Encode( out.bytes );
    }

    std::string Metadata::FormatMetadataNodeName( const Metadata* metadata, bool removeGroups )
    {
        std::string text;
        if( metadata->formattedName )
        {
            text = metadata->formattedName;
        }
        else
        {
            // "DomainWarpGradient" -> "Domain Warp Gradient"
            for( const char* c = metadata->name; *c; c++ )
            {
                bool boundary = !text.empty() && std::islower( (unsigned char)text.back() ) && ( std::isupper( (unsigned char)*c ) || std::isdigit( (unsigned char)*c ) );
                if( boundary ) text += ' ';
                text += *c;
            }
        }

        if( removeGroups )
        {
            for( const char* group : metadata->groups )
            {
                std::size_t at = text.find( group );
                if( at != std::string::npos ) text.erase( at, std::strlen( group ) + 1 );
            }
        }
        return text.empty() ? std::string( metadata->name ) : text;
    }

    std::string Metadata::FormatMetadataMemberName( const Member& member )
    {
        if( member.dimensionIdx < 0 ) return member.name;
        return std::string( kDim_Strings[member.dimensionIdx] ) + " " + member.name;
    }
}
