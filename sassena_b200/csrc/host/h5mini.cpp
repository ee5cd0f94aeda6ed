// h5mini.cpp — see h5mini.hpp.  Everything is little-endian; "addresses" are byte offsets from the superblock's base
// address; sizes of offsets and lengths are 8 bytes throughout.
#include "h5mini.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>

#include "sassena_host.hpp"

namespace sassena {
namespace h5 {

namespace {

const uint8_t kSignature[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
constexpr uint64_t UNDEF = ~(uint64_t)0;
constexpr int kGroupLeafK = 4;       // a symbol node holds up to 2K entries
constexpr int kGroupInternalK = 16;  // group B-tree nodes are sized for 2K children
constexpr int kChunkK = 32;          // "indexed storage internal node K" implied by a version-0 superblock

enum MsgType : uint16_t {
    MSG_NIL = 0x0,
    MSG_DATASPACE = 0x1,
    MSG_DATATYPE = 0x3,
    MSG_FILL_OLD = 0x4,
    MSG_FILL = 0x5,
    MSG_LAYOUT = 0x8,
    MSG_FILTER = 0xB,
    MSG_CONTINUATION = 0x10,
    MSG_SYMBOL_TABLE = 0x11
};

// ------------------------------------------------------------------------------------------------ writer
struct Image {
    std::vector<uint8_t> b;
    uint64_t alloc(size_t n) {
        size_t off = (b.size() + 7) & ~(size_t)7;
        b.resize(off + n, 0);
        return off;
    }
    void put(uint64_t off, uint64_t v, int nbytes) {
        for (int i = 0; i < nbytes; i++) b[off + i] = (uint8_t)(v >> (8 * i));
    }
    void put_bytes(uint64_t off, const void *p, size_t n) {
        if (n) memcpy(&b[off], p, n);
    }
};

struct Bytes {  // little-endian byte builder for message bodies
    std::vector<uint8_t> v;
    Bytes &u(uint64_t x, int n) {
        for (int i = 0; i < n; i++) v.push_back((uint8_t)(x >> (8 * i)));
        return *this;
    }
    Bytes &pad8() {
        while (v.size() % 8) v.push_back(0);
        return *this;
    }
};

struct Message {
    uint16_t type;
    uint8_t flags;
    Bytes body;
};

// version-1 object header: 12-byte prefix padded to 16, then messages (8-byte header + body, bodies padded to 8)
uint64_t write_object_header(Image &img, std::vector<Message> &msgs) {
    size_t total = 0;
    for (auto &m : msgs) {
        m.body.pad8();
        total += 8 + m.body.v.size();
    }
    const uint64_t at = img.alloc(16 + total);
    img.put(at, 1, 1);  // version
    img.put(at + 2, msgs.size(), 2);
    img.put(at + 4, 1, 4);  // object reference count
    img.put(at + 8, total, 4);
    uint64_t p = at + 16;
    for (auto &m : msgs) {
        img.put(p, m.type, 2);
        img.put(p + 2, m.body.v.size(), 2);
        img.put(p + 4, m.flags, 1);
        img.put_bytes(p + 8, m.body.v.data(), m.body.v.size());
        p += 8 + m.body.v.size();
    }
    return at;
}

struct ChunkRef {
    std::vector<uint64_t> offset;  // rank + 1 entries, the last is 0
    uint32_t nbytes;
    uint64_t addr;
};

// v1 B-tree of raw data chunks (node type 1).  Entries arrive in lexicographic offset order.
uint64_t write_chunk_btree(Image &img, std::vector<ChunkRef> entries, const std::vector<uint64_t> &chunk_dims_plus) {
    const size_t nd = chunk_dims_plus.size();  // rank + 1
    const size_t keysize = 8 + 8 * nd;
    const size_t nodesize = 24 + 2 * kChunkK * 8 + (2 * kChunkK + 1) * keysize;
    // the key closing the whole tree: one chunk beyond the last one in every dimension
    ChunkRef last;
    last.nbytes = 0;
    last.addr = UNDEF;
    last.offset.assign(nd, 0);
    if (!entries.empty())
        for (size_t d = 0; d < nd; d++) last.offset[d] = entries.back().offset[d] + chunk_dims_plus[d];
    int level = 0;
    while (true) {
        const size_t per = 2 * kChunkK;
        const size_t nnodes = std::max<size_t>(1, (entries.size() + per - 1) / per);
        std::vector<uint64_t> addr(nnodes);
        for (size_t n = 0; n < nnodes; n++) addr[n] = img.alloc(nodesize);
        std::vector<ChunkRef> parents;
        for (size_t n = 0; n < nnodes; n++) {
            const size_t e0 = n * per, e1 = std::min(entries.size(), e0 + per);
            const uint64_t at = addr[n];
            img.put_bytes(at, "TREE", 4);
            img.put(at + 4, 1, 1);
            img.put(at + 5, level, 1);
            img.put(at + 6, e1 - e0, 2);
            img.put(at + 8, n > 0 ? addr[n - 1] : UNDEF, 8);
            img.put(at + 16, n + 1 < nnodes ? addr[n + 1] : UNDEF, 8);
            uint64_t p = at + 24;
            auto put_key = [&](const ChunkRef &k) {
                img.put(p, k.nbytes, 4);
                img.put(p + 4, 0, 4);  // filter mask
                for (size_t d = 0; d < nd; d++) img.put(p + 8 + 8 * d, k.offset[d], 8);
                p += keysize;
            };
            for (size_t e = e0; e < e1; e++) {
                put_key(entries[e]);
                img.put(p, entries[e].addr, 8);
                p += 8;
            }
            if (e1 > e0) put_key(e1 < entries.size() ? entries[e1] : last);  // right key = left key of the next node
            if (e1 > e0) {
                ChunkRef up = entries[e0];
                up.addr = at;
                parents.push_back(up);
            }
        }
        if (nnodes == 1) return addr[0];
        entries.swap(parents);
        level++;
    }
}

uint64_t product(const std::vector<uint64_t> &v) {
    uint64_t p = 1;
    for (uint64_t x : v) p *= x;
    return p;
}

uint64_t write_dataset(Image &img, const Dataset &ds) {
    std::vector<uint64_t> dims = ds.dims;
    size_t elem = 8;
    const uint8_t *raw = nullptr;
    std::string text;
    Bytes dtype;
    if (ds.kind == Dataset::F64) {
        if (ds.f64.size() != product(dims)) throw Error("h5: dataset " + ds.name + ": data does not match its extent");
        raw = reinterpret_cast<const uint8_t *>(ds.f64.data());
        // class 1 (floating point) version 1; little-endian, implied-msb mantissa, sign at bit 63; IEEE binary64 fields
        dtype.u(0x11, 1).u(0x20, 1).u(0x3f, 1).u(0, 1).u(8, 4);
        dtype.u(0, 2).u(64, 2).u(52, 1).u(11, 1).u(0, 1).u(52, 1).u(1023, 4);
    } else if (ds.kind == Dataset::CHAR) {
        elem = 1;
        dims.assign(1, ds.bytes.size());
        raw = reinterpret_cast<const uint8_t *>(ds.bytes.data());
        dtype.u(0x10, 1).u(0x08, 1).u(0, 1).u(0, 1).u(1, 4).u(0, 2).u(8, 2);  // class 0, signed, 8 bits
    } else {
        text = ds.bytes;
        text.push_back('\0');
        elem = text.size();
        dims.clear();  // scalar
        raw = reinterpret_cast<const uint8_t *>(text.data());
        dtype.u(0x13, 1).u(0x00, 1).u(0, 1).u(0, 1).u(elem, 4);  // class 3, null-terminated, ASCII
    }
    const size_t rank = dims.size();
    const bool has_max = !ds.maxdims.empty();
    if (has_max && ds.maxdims.size() != rank) throw Error("h5: dataset " + ds.name + ": maxdims rank mismatch");
    const bool chunked = !ds.chunk.empty();
    if (chunked && ds.chunk.size() != rank) throw Error("h5: dataset " + ds.name + ": chunk rank mismatch");
    if (has_max && !chunked)
        for (size_t d = 0; d < rank; d++)
            if (ds.maxdims[d] != dims[d]) throw Error("h5: dataset " + ds.name + ": an extendible dataset needs chunks");

    Message space{MSG_DATASPACE, 0, {}};
    space.body.u(1, 1).u(rank, 1).u(has_max ? 1 : 0, 1).u(0, 5);
    for (uint64_t d : dims) space.body.u(d, 8);
    if (has_max)
        for (uint64_t d : ds.maxdims) space.body.u(d, 8);
    Message type{MSG_DATATYPE, 1, dtype};
    Message fill{MSG_FILL, 1, {}};
    if (ds.kind == Dataset::F64)  // fill value 0.0 as the reference sets it; allocation incremental (chunked) / late
        fill.body.u(2, 1).u(chunked ? 3 : 2, 1).u(2, 1).u(1, 1).u(8, 4).u(0, 8);
    else
        fill.body.u(2, 1).u(2, 1).u(2, 1).u(0, 1);
    Message layout{MSG_LAYOUT, 0, {}};
    const uint64_t nbytes = product(dims) * elem;
    if (!chunked) {
        uint64_t addr = UNDEF;
        if (nbytes) {
            addr = img.alloc(nbytes);
            img.put_bytes(addr, raw, nbytes);
        }
        layout.body.u(3, 1).u(1, 1).u(addr, 8).u(nbytes, 8);
    } else {
        std::vector<uint64_t> cd = ds.chunk;
        for (uint64_t c : cd)
            if (c == 0) throw Error("h5: dataset " + ds.name + ": zero chunk dimension");
        const uint64_t chunk_bytes = product(cd) * elem;
        if (chunk_bytes >= ((uint64_t)1 << 32)) throw Error("h5: dataset " + ds.name + ": chunk larger than 4 GiB");
        std::vector<uint64_t> grid(rank), idx(rank, 0);
        uint64_t nchunks = 1;
        for (size_t d = 0; d < rank; d++) {
            grid[d] = (dims[d] + cd[d] - 1) / cd[d];
            nchunks *= grid[d];
        }
        std::vector<uint64_t> stride(rank, 1), cstride(rank, 1);  // element strides of the dataset / of a chunk
        for (size_t d = rank; d-- > 1;) {
            stride[d - 1] = stride[d] * dims[d];
            cstride[d - 1] = cstride[d] * cd[d];
        }
        std::vector<ChunkRef> refs;
        for (uint64_t c = 0; c < nchunks; c++) {
            ChunkRef r;
            r.offset.assign(rank + 1, 0);
            for (size_t d = 0; d < rank; d++) r.offset[d] = idx[d] * cd[d];
            r.nbytes = (uint32_t)chunk_bytes;
            r.addr = img.alloc(chunk_bytes);
            // copy the part of the chunk inside the extent, one innermost run at a time; the rest stays fill value (0)
            std::vector<uint64_t> in(rank, 0), ext(rank);
            for (size_t d = 0; d < rank; d++) ext[d] = std::min(cd[d], dims[d] - r.offset[d]);
            const uint64_t run = ext[rank - 1];
            while (true) {
                uint64_t src = 0, dst = 0;
                for (size_t d = 0; d < rank; d++) {
                    src += (r.offset[d] + in[d]) * stride[d];
                    dst += in[d] * cstride[d];
                }
                img.put_bytes(r.addr + dst * elem, raw + src * elem, run * elem);
                size_t d = rank - 1;
                while (d-- > 0) {
                    if (++in[d] < ext[d]) break;
                    in[d] = 0;
                }
                if (d == (size_t)-1) break;
            }
            refs.push_back(r);
            for (size_t d = rank; d-- > 0;) {  // next chunk, row-major (= lexicographic offsets)
                if (++idx[d] < grid[d]) break;
                idx[d] = 0;
            }
        }
        std::vector<uint64_t> cdp = cd;
        cdp.push_back(elem);
        const uint64_t bt = write_chunk_btree(img, refs, cdp);
        layout.body.u(3, 1).u(2, 1).u(rank + 1, 1).u(bt, 8);
        for (uint64_t c : cdp) layout.body.u(c, 4);
    }
    std::vector<Message> msgs{space, type, fill, layout};
    // spare room (a NIL message) so that a library appending e.g. a modification time does not need a continuation block
    Message nil{MSG_NIL, 0, {}};
    nil.body.u(0, 8).u(0, 8).u(0, 8);
    msgs.push_back(nil);
    return write_object_header(img, msgs);
}

struct GroupRef {
    uint64_t header, btree, heap;
};

GroupRef write_group(Image &img, const Group &g) {
    struct Child {
        std::string name;
        uint64_t header;
        bool is_group;
        GroupRef ref;
    };
    std::vector<Child> kids;
    for (auto &sub : g.groups) {
        GroupRef r = write_group(img, sub);
        kids.push_back({sub.name, r.header, true, r});
    }
    for (auto &ds : g.datasets) kids.push_back({ds.name, write_dataset(img, ds), false, {}});
    std::sort(kids.begin(), kids.end(), [](const Child &a, const Child &b) { return a.name < b.name; });
    for (size_t i = 0; i < kids.size(); i++) {
        if (kids[i].name.empty() || kids[i].name.find('/') != std::string::npos) throw Error("h5: bad link name");
        if (i && kids[i].name == kids[i - 1].name) throw Error("h5: duplicate link name " + kids[i].name);
    }
    if (kids.size() > 2 * kGroupLeafK) throw Error("h5: more than 8 links in one group are not supported by this writer");
    // local heap: the empty string at offset 0, then the names, then one free block
    std::vector<uint64_t> name_off(kids.size());
    uint64_t used = 8;
    for (size_t i = 0; i < kids.size(); i++) {
        name_off[i] = used;
        used += (kids[i].name.size() + 1 + 7) & ~(size_t)7;
    }
    const uint64_t seg = std::max<uint64_t>(used + 32, 88);
    const uint64_t heap = img.alloc(32 + seg);
    img.put_bytes(heap, "HEAP", 4);
    img.put(heap + 8, seg, 8);
    img.put(heap + 16, used, 8);       // head of the free list
    img.put(heap + 24, heap + 32, 8);  // data segment address
    for (size_t i = 0; i < kids.size(); i++) img.put_bytes(heap + 32 + name_off[i], kids[i].name.data(), kids[i].name.size());
    img.put(heap + 32 + used, 1, 8);  // free block: no next block (H5HL_FREE_NULL)
    img.put(heap + 32 + used + 8, seg - used, 8);
    // symbol node
    uint64_t snod = UNDEF;
    if (!kids.empty()) {
        snod = img.alloc(8 + 2 * kGroupLeafK * 40);
        img.put_bytes(snod, "SNOD", 4);
        img.put(snod + 4, 1, 1);
        img.put(snod + 6, kids.size(), 2);
        for (size_t i = 0; i < kids.size(); i++) {
            const uint64_t e = snod + 8 + 40 * i;
            img.put(e, name_off[i], 8);
            img.put(e + 8, kids[i].header, 8);
            if (kids[i].is_group) {  // cached symbol-table information
                img.put(e + 16, 1, 4);
                img.put(e + 24, kids[i].ref.btree, 8);
                img.put(e + 32, kids[i].ref.heap, 8);
            }
        }
    }
    // group B-tree (node type 0): one leaf pointing at the symbol node
    const uint64_t bt = img.alloc(24 + (2 * kGroupInternalK + 1) * 8 + 2 * kGroupInternalK * 8);
    img.put_bytes(bt, "TREE", 4);
    img.put(bt + 6, kids.empty() ? 0 : 1, 2);
    img.put(bt + 8, UNDEF, 8);
    img.put(bt + 16, UNDEF, 8);
    if (!kids.empty()) {
        img.put(bt + 24, 0, 8);                 // key 0: the empty string
        img.put(bt + 32, snod, 8);              // child 0
        img.put(bt + 40, name_off.back(), 8);   // key 1: the largest name in child 0
    }
    Message st{MSG_SYMBOL_TABLE, 0, {}};
    st.body.u(bt, 8).u(heap, 8);
    std::vector<Message> msgs{st};
    return {write_object_header(img, msgs), bt, heap};
}

}  // namespace

std::vector<uint8_t> serialize(const Group &root) {
    Image img;
    img.alloc(96);  // superblock
    const GroupRef r = write_group(img, root);
    img.put_bytes(0, kSignature, 8);
    // versions: superblock 0, free-space 0, root symbol table entry 0, (reserved), shared header messages 0
    img.put(13, 8, 1);  // size of offsets
    img.put(14, 8, 1);  // size of lengths
    img.put(16, kGroupLeafK, 2);
    img.put(18, kGroupInternalK, 2);
    img.put(20, 0, 4);      // consistency flags
    img.put(24, 0, 8);      // base address
    img.put(32, UNDEF, 8);  // free-space info
    img.put(40, img.b.size(), 8);  // end of file
    img.put(48, UNDEF, 8);  // driver info
    img.put(56, 0, 8);      // root entry: link name offset
    img.put(64, r.header, 8);
    img.put(72, 1, 4);  // cache type 1: symbol table
    img.put(80, r.btree, 8);
    img.put(88, r.heap, 8);
    return img.b;
}

void write_file(const std::string &path, const Group &root) {
    const std::vector<uint8_t> b = serialize(root);
    const std::string tmp = path + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) throw Error("cannot create " + tmp);
    const bool ok = fwrite(b.data(), 1, b.size(), f) == b.size();
    if (fclose(f) != 0 || !ok) throw Error("short write to " + tmp);
    if (rename(tmp.c_str(), path.c_str()) != 0) throw Error("cannot rename " + tmp + " to " + path);
}

const Dataset *Group::find(const std::string &n) const {
    for (auto &d : datasets)
        if (d.name == n) return &d;
    return nullptr;
}
const Group *Group::find_group(const std::string &n) const {
    for (auto &g : groups)
        if (g.name == n) return &g;
    return nullptr;
}

// ------------------------------------------------------------------------------------------------ reader
namespace {

struct Reader {
    const std::vector<uint8_t> &b;
    uint64_t base = 0;
    explicit Reader(const std::vector<uint8_t> &img) : b(img) {}

    uint64_t get(uint64_t off, int n) const {  // absolute file offset
        if (off > b.size() || (uint64_t)n > b.size() - off) throw Error("h5: read beyond the end of the file");
        uint64_t v = 0;
        for (int i = 0; i < n; i++) v |= (uint64_t)b[off + i] << (8 * i);
        return v;
    }
    uint64_t at(uint64_t addr) const { return base + addr; }  // file address -> absolute offset
    // [off, off + n) lies inside the file (written so that a corrupt address near 2^64 cannot wrap around the test)
    bool inside(uint64_t off, uint64_t n) const { return off <= b.size() && n <= b.size() - off; }
    void expect(uint64_t off, const char *sig) const {
        if (!inside(off, 4) || memcmp(&b[off], sig, 4) != 0) throw Error(std::string("h5: missing signature ") + sig);
    }
    // bytes of an extent; a corrupt dataspace must not turn into an allocation of petabytes: chunked datasets may be
    // sparser than their extent, so the bound is generous (64 x the file + 64 MiB), overflow is refused outright
    uint64_t extent_bytes(const std::vector<uint64_t> &dims, uint64_t elem, const std::string &name) const {
        const uint64_t cap = 64 * (uint64_t)b.size() + ((uint64_t)64 << 20);
        uint64_t p = elem;
        for (uint64_t x : dims) {
            if (x != 0 && p > cap / x) throw Error("h5: dataset " + name + ": extent not plausible for a file of this size");
            p *= x;
        }
        if (p > cap) throw Error("h5: dataset " + name + ": extent not plausible for a file of this size");
        return p;
    }

    struct Msg {
        uint16_t type;
        uint64_t off;  // absolute offset of the body
        uint16_t size;
    };
    std::vector<Msg> object_header(uint64_t addr) const {
        const uint64_t h = at(addr);
        if (get(h, 1) != 1) throw Error("h5: only version-1 object headers are supported");
        size_t nmsg = get(h + 2, 2);
        std::vector<std::pair<uint64_t, uint64_t>> blocks{{h + 16, get(h + 8, 4)}};
        std::vector<Msg> out;
        for (size_t bi = 0; bi < blocks.size(); bi++) {
            uint64_t p = blocks[bi].first;
            const uint64_t end = p + blocks[bi].second;
            while (p + 8 <= end && out.size() < nmsg) {
                Msg m{(uint16_t)get(p, 2), p + 8, (uint16_t)get(p + 2, 2)};
                if (m.off + m.size > end) throw Error("h5: object header message overruns its block");
                if (m.type == MSG_CONTINUATION) blocks.push_back({at(get(m.off, 8)), get(m.off + 8, 8)});
                out.push_back(m);
                p = m.off + m.size;
            }
        }
        return out;
    }

    std::string heap_string(uint64_t heap_addr, uint64_t offset) const {
        const uint64_t h = at(heap_addr);
        expect(h, "HEAP");
        const uint64_t seg_size = get(h + 8, 8), seg = at(get(h + 24, 8));
        std::string s;
        for (uint64_t i = offset; i < seg_size; i++) {
            const char c = (char)get(seg + i, 1);
            if (!c) break;
            s.push_back(c);
        }
        return s;
    }

    // symbol table entries of a group B-tree, in order
    // (a child sits exactly one level below its parent: a corrupt tree that points back into itself is refused instead of
    // being walked 2^depth times)
    void group_entries(uint64_t bt_addr, uint64_t heap, std::vector<std::pair<std::string, uint64_t>> &out, int depth = 0,
                       int want_level = -1) const {
        if (depth > 32) throw Error("h5: group B-tree too deep");
        const uint64_t n = at(bt_addr);
        expect(n, "TREE");
        if (get(n + 4, 1) != 0) throw Error("h5: group B-tree node of the wrong type");
        const int level = (int)get(n + 5, 1);
        if (want_level >= 0 && level != want_level) throw Error("h5: group B-tree levels are inconsistent");
        const size_t used = get(n + 6, 2);
        for (size_t i = 0; i < used; i++) {
            const uint64_t child = get(n + 24 + 8 + 16 * i, 8);
            if (level > 0) {
                group_entries(child, heap, out, depth + 1, level - 1);
                continue;
            }
            const uint64_t s = at(child);
            expect(s, "SNOD");
            const size_t nsym = get(s + 6, 2);
            for (size_t k = 0; k < nsym; k++) {
                const uint64_t e = s + 8 + 40 * k;
                out.push_back({heap_string(heap, get(e, 8)), get(e + 8, 8)});
            }
        }
    }

    struct ChunkLoc {
        std::vector<uint64_t> offset;
        uint64_t nbytes, addr;
    };
    void chunk_entries(uint64_t bt_addr, size_t nd, std::vector<ChunkLoc> &out, int depth = 0, int want_level = -1) const {
        if (depth > 32) throw Error("h5: chunk B-tree too deep");
        const uint64_t n = at(bt_addr);
        expect(n, "TREE");
        if (get(n + 4, 1) != 1) throw Error("h5: chunk B-tree node of the wrong type");
        const int level = (int)get(n + 5, 1);
        if (want_level >= 0 && level != want_level) throw Error("h5: chunk B-tree levels are inconsistent");
        const size_t used = get(n + 6, 2);
        const size_t keysize = 8 + 8 * nd;
        for (size_t i = 0; i < used; i++) {
            const uint64_t k = n + 24 + i * (keysize + 8);
            const uint64_t child = get(k + keysize, 8);
            if (level > 0) {
                chunk_entries(child, nd, out, depth + 1, level - 1);
                continue;
            }
            if (get(k + 4, 4) != 0) throw Error("h5: filtered chunks are not supported");
            ChunkLoc c;
            c.nbytes = get(k, 4);
            c.addr = child;
            for (size_t d = 0; d < nd; d++) c.offset.push_back(get(k + 8 + 8 * d, 8));
            out.push_back(c);
        }
    }

    Dataset dataset(const std::string &name, const std::vector<Msg> &msgs) const {
        Dataset ds;
        ds.name = name;
        size_t elem = 0;
        bool have_space = false, have_type = false, have_layout = false;
        Msg layout{};
        for (auto &m : msgs) {
            if (m.type == MSG_DATASPACE) {
                const int ver = (int)get(m.off, 1);
                const size_t rank = get(m.off + 1, 1);
                const int flags = (int)get(m.off + 2, 1);
                if (ver != 1 && ver != 2) throw Error("h5: dataspace version not supported");
                uint64_t p = m.off + (ver == 1 ? 8 : 4);
                for (size_t d = 0; d < rank; d++, p += 8) ds.dims.push_back(get(p, 8));
                if (flags & 1)
                    for (size_t d = 0; d < rank; d++, p += 8) ds.maxdims.push_back(get(p, 8));
                have_space = true;
            } else if (m.type == MSG_DATATYPE) {
                const int cls = (int)get(m.off, 1) & 0x0f;
                elem = get(m.off + 4, 4);
                if (cls == 1 && elem == 8 && (get(m.off + 1, 1) & 1) == 0)
                    ds.kind = Dataset::F64;
                else if (cls == 0 && elem == 1)
                    ds.kind = Dataset::CHAR;
                else if (cls == 3)
                    ds.kind = Dataset::STRING;
                else
                    throw Error("h5: dataset " + name + ": datatype not supported (only f64 LE, 1-byte integers, fixed strings)");
                have_type = true;
            } else if (m.type == MSG_LAYOUT) {
                layout = m;
                have_layout = true;
            } else if (m.type == MSG_FILTER) {
                throw Error("h5: dataset " + name + ": filters are not supported");
            }
        }
        if (!have_space || !have_type || !have_layout) throw Error("h5: dataset " + name + ": incomplete object header");
        const size_t rank = ds.dims.size();
        const uint64_t total = extent_bytes(ds.dims, elem, name);
        std::vector<uint8_t> raw(total, 0);
        const int ver = (int)get(layout.off, 1);
        int cls;
        uint64_t addr = UNDEF;
        std::vector<uint64_t> ldims;
        if (ver == 1 || ver == 2) {
            const size_t nd = get(layout.off + 1, 1);
            cls = (int)get(layout.off + 2, 1);
            uint64_t p = layout.off + 8;
            if (cls != 0) {
                addr = get(p, 8);
                p += 8;
            }
            for (size_t d = 0; d < nd; d++, p += 4) ldims.push_back(get(p, 4));
            if (cls == 0) {  // compact: size + data follow
                const uint64_t sz = get(p, 4);
                if (sz < total) throw Error("h5: compact dataset shorter than its extent");
                for (uint64_t i = 0; i < total; i++) raw[i] = (uint8_t)get(p + 4 + i, 1);
            }
        } else if (ver == 3) {
            cls = (int)get(layout.off + 1, 1);
            if (cls == 0) {
                const uint64_t sz = get(layout.off + 2, 2);
                if (sz < total) throw Error("h5: compact dataset shorter than its extent");
                for (uint64_t i = 0; i < total; i++) raw[i] = (uint8_t)get(layout.off + 4 + i, 1);
            } else if (cls == 1) {
                addr = get(layout.off + 2, 8);
            } else {
                const size_t nd = get(layout.off + 2, 1);
                addr = get(layout.off + 3, 8);
                for (size_t d = 0; d < nd; d++) ldims.push_back(get(layout.off + 11 + 4 * d, 4));
            }
        } else {
            throw Error("h5: data layout version not supported");
        }
        if (cls == 1) {
            if (addr != UNDEF && total) {
                if (!inside(at(addr), total)) throw Error("h5: dataset " + name + " extends beyond the file");
                memcpy(raw.data(), &b[at(addr)], total);
            }
        } else if (cls == 2) {
            if (ldims.size() != rank + 1) throw Error("h5: chunk rank mismatch in " + name);
            ds.chunk.assign(ldims.begin(), ldims.begin() + rank);
            std::vector<ChunkLoc> chunks;
            if (addr != UNDEF) chunk_entries(addr, rank + 1, chunks);
            std::vector<uint64_t> stride(rank, 1), cstride(rank, 1);
            for (size_t d = rank; d-- > 1;) {
                stride[d - 1] = stride[d] * ds.dims[d];
                cstride[d - 1] = cstride[d] * ds.chunk[d];
            }
            const uint64_t chunk_bytes = extent_bytes(ds.chunk, elem, name);
            for (auto &c : chunks) {
                if (c.nbytes != chunk_bytes) throw Error("h5: chunk of unexpected size in " + name);
                if (!inside(at(c.addr), chunk_bytes)) throw Error("h5: chunk beyond the end of the file");
                bool inside = rank > 0;
                for (size_t d = 0; d < rank; d++) inside = inside && c.offset[d] < ds.dims[d];
                if (!inside) continue;  // a chunk left over from a larger extent
                std::vector<uint64_t> in(rank, 0), ext(rank);
                for (size_t d = 0; d < rank; d++) ext[d] = std::min(ds.chunk[d], ds.dims[d] - c.offset[d]);
                const uint64_t run = ext[rank - 1];
                while (true) {
                    uint64_t dst = 0, src = 0;
                    for (size_t d = 0; d < rank; d++) {
                        dst += (c.offset[d] + in[d]) * stride[d];
                        src += in[d] * cstride[d];
                    }
                    memcpy(&raw[dst * elem], &b[at(c.addr) + src * elem], run * elem);
                    size_t d = rank - 1;
                    while (d-- > 0) {
                        if (++in[d] < ext[d]) break;
                        in[d] = 0;
                    }
                    if (d == (size_t)-1) break;
                }
            }
        }
        if (ds.kind == Dataset::F64) {
            ds.f64.resize(total / 8);
            if (total) memcpy(ds.f64.data(), raw.data(), total);
        } else if (ds.kind == Dataset::CHAR) {
            ds.bytes.assign(raw.begin(), raw.end());
        } else {
            ds.bytes.assign(raw.begin(), raw.end());
            const size_t z = ds.bytes.find('\0');
            if (z != std::string::npos) ds.bytes.resize(z);
        }
        return ds;
    }

    Group group(const std::string &name, uint64_t bt, uint64_t heap, int depth) const {
        if (depth > 16) throw Error("h5: groups nested too deeply");
        if (++objects_ > 100000) throw Error("h5: too many objects (a group that contains itself?)");
        Group g;
        g.name = name;
        std::vector<std::pair<std::string, uint64_t>> entries;
        group_entries(bt, heap, entries);
        for (auto &e : entries) {
            const std::vector<Msg> msgs = object_header(e.second);
            const Msg *st = nullptr;
            for (auto &m : msgs)
                if (m.type == MSG_SYMBOL_TABLE) st = &m;
            if (st)
                g.groups.push_back(group(e.first, get(st->off, 8), get(st->off + 8, 8), depth + 1));
            else if (++objects_ > 100000)
                throw Error("h5: too many objects (a group that contains itself?)");
            else
                g.datasets.push_back(dataset(e.first, msgs));
        }
        return g;
    }
    mutable size_t objects_ = 0;
};

}  // namespace

Group parse(const std::vector<uint8_t> &image) {
    Reader r(image);
    // the superblock sits at offset 0 or at 512, 1024, 2048, ... (after a user block)
    uint64_t sb = UNDEF;
    for (uint64_t off = 0; off + 8 <= image.size(); off = off ? off * 2 : 512)
        if (memcmp(&image[off], kSignature, 8) == 0) {
            sb = off;
            break;
        }
    if (sb == UNDEF) throw Error("h5: not an HDF5 file (no superblock signature)");
    const int ver = (int)r.get(sb + 8, 1);
    if (ver != 0 && ver != 1) throw Error("h5: only superblock versions 0 and 1 are supported");
    if (r.get(sb + 13, 1) != 8 || r.get(sb + 14, 1) != 8) throw Error("h5: only 8-byte offsets and lengths are supported");
    const uint64_t p = sb + (ver == 0 ? 24 : 28);
    // addresses are relative to the base address; with a user block the library stores base = its size
    r.base = r.get(p, 8);
    const uint64_t root = p + 32;  // root group symbol table entry
    const uint64_t header = r.get(root + 8, 8);
    uint64_t bt, heap;
    if (r.get(root + 16, 4) == 1) {
        bt = r.get(root + 24, 8);
        heap = r.get(root + 32, 8);
    } else {
        const auto msgs = r.object_header(header);
        const Reader::Msg *st = nullptr;
        for (auto &m : msgs)
            if (m.type == MSG_SYMBOL_TABLE) st = &m;
        if (!st) throw Error("h5: the root group has no symbol table");
        bt = r.get(st->off, 8);
        heap = r.get(st->off + 8, 8);
    }
    return r.group("/", bt, heap, 0);
}

Group read_file(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) throw Error("cannot open " + path);
    std::vector<uint8_t> b;
    uint8_t buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) b.insert(b.end(), buf, buf + n);
    fclose(f);
    return parse(b);
}

}  // namespace h5

// ------------------------------------------------------------------------------------------------ signal file
SignalFileH5::SignalFileH5(const std::string &path, size_t NF, size_t chunksize, bool fqt, bool fq0, bool fq, bool fq2)
    : path_(path), NF_(NF), chunksize_(chunksize ? chunksize : 1), fqt_(fqt), fq0_(fq0), fq_(fq), fq2_(fq2) {
    if (NF < 1) throw Error("Number of frames must not be negative or zero!");  // file_writer_service.cpp:23-26
}

void SignalFileH5::set_meta(const std::string &rawconfig, const std::string &config, const std::string &database) {
    rawconfig_ = rawconfig;
    config_ = config;
    database_ = database;
}

std::vector<double> SignalFileH5::init() {
    FILE *f = fopen(path_.c_str(), "rb");
    if (!f) {  // init_new (file_writer_service.cpp:44-171)
        flush();
        return {};
    }
    fclose(f);
    h5::Group old;
    try {
        old = h5::read_file(path_);
    } catch (const Error &e) {
        throw Error(path_ + " does not to be a HDF5 data file. aborting... (" + e.what() + ")");  // :28-36
    }
    const h5::Dataset *q = old.find("qvectors");
    if (!q || q->dims.size() != 2 || q->dims[1] != 3) throw Error(path_ + ": no qvectors dataset");
    const size_t n = q->dims[0];
    auto take = [&](const char *name, bool wanted, size_t width, std::vector<double> &dst) {
        if (!wanted) return;
        const h5::Dataset *d = old.find(name);
        if (!d || d->dims.empty() || d->dims[0] != n || d->f64.size() != n * width)
            throw Error(path_ + ": dataset " + name + " does not match qvectors / the number of frames");  // test_fqt_dim
        dst = d->f64;
    };
    take("fqt", fqt_, 2 * NF_, vfqt_);
    take("fq0", fq0_, 2, vfq0_);
    take("fq", fq_, 2, vfq_);
    take("fq2", fq2_, 2, vfq2_);
    q_ = q->f64;
    return q_;
}

void SignalFileH5::write(const double q[3], const double *fqt, const double fq[2], const double fq2[2]) {
    q_.insert(q_.end(), q, q + 3);
    if (fqt_) vfqt_.insert(vfqt_.end(), fqt, fqt + 2 * NF_);
    if (fq0_) vfq0_.insert(vfq0_.end(), fqt, fqt + 2);  // fq0 = fqt[0] (file_writer_service.cpp:514)
    if (fq_) vfq_.insert(vfq_.end(), fq, fq + 2);
    if (fq2_) vfq2_.insert(vfq2_.end(), fq2, fq2 + 2);
}

void SignalFileH5::flush() const {
    const uint64_t n = q_.size() / 3;
    h5::Group root, meta;
    meta.name = "meta";
    h5::Dataset raw, cfg, db;
    raw.name = "rawconfig";
    raw.kind = h5::Dataset::CHAR;
    raw.bytes = rawconfig_;
    cfg.name = "config";
    cfg.kind = h5::Dataset::STRING;
    cfg.bytes = config_;
    db.name = "database";
    db.kind = h5::Dataset::STRING;
    db.bytes = database_;
    meta.datasets = {raw, cfg, db};
    root.groups.push_back(meta);
    auto rows = [&](const char *name, const std::vector<double> &v, std::vector<uint64_t> tail, std::vector<uint64_t> chunk) {
        h5::Dataset d;
        d.name = name;
        d.dims = {n};
        d.maxdims = {h5::UNLIMITED};
        for (uint64_t t : tail) {
            d.dims.push_back(t);
            d.maxdims.push_back(t);
        }
        d.chunk = chunk;
        d.f64 = v;
        root.datasets.push_back(d);
    };
    rows("qvectors", q_, {3}, {chunksize_, 3});
    if (fqt_) {  // chunk shape of file_writer_service.cpp:86-92
        const uint64_t c2 = (NF_ > 0 && NF_ < chunksize_) ? NF_ : chunksize_;
        const uint64_t c1 = std::max<uint64_t>(1, chunksize_ / c2);
        rows("fqt", vfqt_, {NF_, 2}, {c1, c2, 2});
    }
    if (fq0_) rows("fq0", vfq0_, {2}, {chunksize_, 2});
    if (fq_) rows("fq", vfq_, {2}, {chunksize_, 2});
    if (fq2_) rows("fq2", vfq2_, {2}, {chunksize_, 2});
    h5::write_file(path_, root);
}

}  // namespace sassena
