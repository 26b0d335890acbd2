// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may include, link or execute anything under oracle/.
//
// CPU restatement of the third-party `seahash = "4.1"` streaming hasher that the
// reference uses for every checksum (reference Cargo.toml:25; call sites
// src/snapshot/mod.rs:315-317 `checksum_hasher()`, src/snapshot/checksum.rs:38-44,
// src/snapshot/component_checksum.rs:44-48,77-93, src/snapshot/entity_checksum.rs:35-43,
// src/snapshot/resource_checksum.rs:41-43, examples/stress_tests/particles.rs:107-120,207-222).
//
// seahash is NOT vendored under /root/reference (no Cargo.lock, no vendor dir), so the
// published algorithm is restated here:
//   state (a,b,c,d) = (0x16f11fe89b0d677c, 0xb480a793d8e6c86c,
//                      0x6fe2e5aaf078ebc9, 0x14f994a4c5259381)
//   diffuse(x): x *= P; x ^= (x >> 32) >> (x >> 60); x *= P;   P = 0x6eed0e9da4d94a4f
//   bytes are appended to a stream; every complete little-endian 8-byte word w does
//       t = diffuse(a ^ w); (a,b,c,d) = (b,c,d,t)
//   finish(): if r tail bytes remain (zero-extended LE word w): a = diffuse(a ^ w)
//             result = diffuse(a ^ b ^ c ^ d ^ total_len_bytes)
//   Rust `Hash` for u8/u32/u64/usize appends the value's little-endian bytes.
//
// PINNING: checked in tests/test_oracle_seahash.py against the seahash crate's own
// documented vectors (hash(b"to be or not to be") == 1988685042348123509,
// hash(b"") == 14492805990617963705) and against the derived vectors of SURVEY.md §8c
// (an independent restatement).  The reference itself holds NO numeric checksum golden
// (its only checksum test asserts self-equality, src/snapshot/checksum.rs:108-113).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace oracle {

inline uint64_t sea_diffuse(uint64_t x) {
    const uint64_t P = 0x6eed0e9da4d94a4fULL;
    x *= P;
    x ^= (x >> 32) >> (x >> 60);
    x *= P;
    return x;
}

struct SeaHasher {
    uint64_t a = 0x16f11fe89b0d677cULL;
    uint64_t b = 0xb480a793d8e6c86cULL;
    uint64_t c = 0x6fe2e5aaf078ebc9ULL;
    uint64_t d = 0x14f994a4c5259381ULL;
    uint64_t written = 0;  // bytes consumed as whole words
    uint64_t tail = 0;     // pending (<8) bytes, little endian
    unsigned ntail = 0;

    void push(uint64_t w) {
        uint64_t t = sea_diffuse(a ^ w);
        a = b; b = c; c = d; d = t;
        written += 8;
    }
    void write(const void* src, size_t n) {
        const uint8_t* p = static_cast<const uint8_t*>(src);
        while (n) {
            size_t take = 8 - ntail;
            if (take > n) take = n;
            for (size_t i = 0; i < take; ++i)
                tail |= static_cast<uint64_t>(p[i]) << (8 * (ntail + i));
            ntail += static_cast<unsigned>(take);
            p += take;
            n -= take;
            if (ntail == 8) { push(tail); tail = 0; ntail = 0; }
        }
    }
    void write_u8(uint8_t v) { write(&v, 1); }
    void write_u32(uint32_t v) { uint8_t le[4]; for (int i = 0; i < 4; ++i) le[i] = uint8_t(v >> (8 * i)); write(le, 4); }
    void write_u64(uint64_t v) { uint8_t le[8]; for (int i = 0; i < 8; ++i) le[i] = uint8_t(v >> (8 * i)); write(le, 8); }
    void write_usize(uint64_t v) { write_u64(v); }  // 64-bit target
    uint64_t finish() const {
        uint64_t aa = ntail ? sea_diffuse(a ^ tail) : a;
        return sea_diffuse(aa ^ b ^ c ^ d ^ (written + ntail));
    }
};

inline uint64_t seahash(const void* p, size_t n) {
    SeaHasher h;
    h.write(p, n);
    return h.finish();
}

}  // namespace oracle
