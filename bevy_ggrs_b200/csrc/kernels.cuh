// CUDA kernels of the rollback engine (sm_100a).  HBM-bound integer / f32 streaming work:
// no tensor cores (there is no dense contraction on this path), the design rules that matter are
// coalescing (word-planar columns => every warp access is a contiguous 128..512 B run),
// 16-byte vector loads/stores, enough bytes in flight per SM, and ONE launch per request vector.
//
// Data layout (DESIGN.md "Data layout in HBM"): an *image* is every registered column split into
// 4-byte word planes of `epad` rows each, followed by a 1-byte-per-row alive plane:
//     image = [plane 0 | plane 1 | ... | plane W-1 | alive]      plane p at p*epad*4
// Image 0 is the live world, image s+1 is snapshot slot s of the ring.  A row is one rollback
// entity; its RollbackOrdered index is order_base + row (rollback.rs:66-83).
//
// Reference semantics implemented here (paths relative to the upstream repo):
//   save  : component_snapshot.rs:66-84   (copy every registered column into the frame's snapshot)
//   load  : component_snapshot.rs:95-123  (overwrite live columns from the snapshot; entity.rs:55-99 -> alive plane)
//   cksum : component_checksum.rs:67-108, entity_checksum.rs:29-52 (XOR of per-entity seahashes; count of live rows)
//   systems: examples/stress_tests/particles.rs:272-289 (update_particles, despawn_particles)
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "seahash.cuh"

namespace bgr {

constexpr int kMaxOps = 80;       // == BGR_MAX_REQUESTS
constexpr int kMaxSaves = 40;
constexpr int kMaxPassive = 64;   // word planes that no compiled system touches
constexpr int kAccStride = 8;     // u64 per save: [0..5] column xors, [6] active rows, [7] flags

enum OpKind : uint32_t { OP_SAVE = 0, OP_LOAD = 1, OP_ADVANCE = 2 };
enum OpFlags : uint32_t { OPF_NO_STORE = 1u };  // ring depth 0: checksum only

struct Op {
    uint32_t kind;
    uint32_t image;       // image index the op reads (LOAD) or writes (SAVE)
    uint32_t dt_bits;     // ADVANCE: Time<GgrsTime>::delta_secs as f32 bits (time.rs:63-76)
    uint32_t n_rows;      // rows that exist while this op runs (RollbackOrdered::len())
    uint32_t save_index;  // SAVE: which accumulator row
    uint32_t flags;
    uint32_t call_count;  // ADVANCE: value of the un-rolled-back host counter (test system only)
    uint8_t inputs[4];    // ADVANCE: first 4 player inputs (u8); systems needing more use the stepwise path
};
static_assert(sizeof(Op) == 32, "Op must stay 32 bytes");

enum ProgFlags : uint32_t {
    PF_READ_LIVE = 1u,           // program does not start with LOAD: initial state comes from image 0
    PF_WRITE_LIVE_ACTIVE = 2u,   // program contains LOAD or ADVANCE: final active planes go to image 0
    PF_WRITE_LIVE_PASSIVE = 4u,  // program contains LOAD: final passive planes go to image 0
    PF_CK_T = 8u, PF_CK_T_FINITE = 16u, PF_CK_V = 32u, PF_CK_V_FINITE = 64u,
};

struct ProgramParams {
    uint8_t* arena;
    unsigned long long image_bytes;
    unsigned long long order_base;
    unsigned long long* accum;  // device [kMaxSaves][kAccStride]
    unsigned long long* out;    // host-mapped [kMaxSaves][kAccStride]
    unsigned int* ticket;
    uint32_t epad, words, n_ops, n_saves;
    uint32_t max_rows, live_rows, flags;
    uint32_t t_plane, v_plane, l_plane;     // first word plane of Transform / Velocity / Ttl
    uint32_t ck_t_slot, ck_v_slot;           // accumulator column of each checksummed type
    uint32_t n_passive;
    uint16_t passive[kMaxPassive];
    Op ops[kMaxOps];
};
static_assert(sizeof(ProgramParams) <= 4000, "kernel parameter block must fit 4 KB");

// ---------------------------------------------------------------------------------------------
// vector access helpers: VEC consecutive rows of one word plane = one 4*VEC byte access
// ---------------------------------------------------------------------------------------------
template <int VEC> struct Vec;
template <> struct Vec<1> { using T = uint32_t; };
template <> struct Vec<2> { using T = uint2; };
template <> struct Vec<4> { using T = uint4; };

template <int VEC> __device__ __forceinline__ void vec_load(const uint8_t* p, uint32_t (&r)[VEC]);
template <> __device__ __forceinline__ void vec_load<1>(const uint8_t* p, uint32_t (&r)[1]) { r[0] = __ldcs(reinterpret_cast<const uint32_t*>(p)); }
template <> __device__ __forceinline__ void vec_load<2>(const uint8_t* p, uint32_t (&r)[2]) { uint2 v = __ldcs(reinterpret_cast<const uint2*>(p)); r[0] = v.x; r[1] = v.y; }
template <> __device__ __forceinline__ void vec_load<4>(const uint8_t* p, uint32_t (&r)[4]) { uint4 v = __ldcs(reinterpret_cast<const uint4*>(p)); r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w; }

template <int VEC> __device__ __forceinline__ void vec_store(uint8_t* p, const uint32_t (&r)[VEC]);
template <> __device__ __forceinline__ void vec_store<1>(uint8_t* p, const uint32_t (&r)[1]) { __stcs(reinterpret_cast<uint32_t*>(p), r[0]); }
template <> __device__ __forceinline__ void vec_store<2>(uint8_t* p, const uint32_t (&r)[2]) { __stcs(reinterpret_cast<uint2*>(p), make_uint2(r[0], r[1])); }
template <> __device__ __forceinline__ void vec_store<4>(uint8_t* p, const uint32_t (&r)[4]) { __stcs(reinterpret_cast<uint4*>(p), make_uint4(r[0], r[1], r[2], r[3])); }

// alive plane: VEC bytes packed little-endian into one register
template <int VEC> __device__ __forceinline__ uint32_t alive_load(const uint8_t* p);
template <> __device__ __forceinline__ uint32_t alive_load<1>(const uint8_t* p) { return __ldcs(p); }
template <> __device__ __forceinline__ uint32_t alive_load<2>(const uint8_t* p) { return __ldcs(reinterpret_cast<const unsigned short*>(p)); }
template <> __device__ __forceinline__ uint32_t alive_load<4>(const uint8_t* p) { return __ldcs(reinterpret_cast<const uint32_t*>(p)); }
template <int VEC> __device__ __forceinline__ void alive_store(uint8_t* p, uint32_t a);
template <> __device__ __forceinline__ void alive_store<1>(uint8_t* p, uint32_t a) { __stcs(p, (unsigned char)a); }
template <> __device__ __forceinline__ void alive_store<2>(uint8_t* p, uint32_t a) { __stcs(reinterpret_cast<unsigned short*>(p), (unsigned short)a); }
template <> __device__ __forceinline__ void alive_store<4>(uint8_t* p, uint32_t a) { __stcs(reinterpret_cast<uint32_t*>(p), a); }

__device__ __forceinline__ bool f32_bits_nonfinite(uint32_t b) { return (b & 0x7f800000u) == 0x7f800000u; }

// mask with byte j = 0x01 for every row (row0 + j) < n_rows
template <int VEC> __device__ __forceinline__ uint32_t rows_mask(uint32_t row0, uint32_t n_rows) {
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < VEC; ++j) m |= (row0 + j < n_rows) ? (1u << (8 * j)) : 0u;
    return m;
}

// ---------------------------------------------------------------------------------------------
// update_particles (particles.rs:272-280) for one entity: every mul and add individually rounded
// (Rust/glam scalar Vec3, no FMA contraction); gravity = Vec3::NEG_Y * 200.0 = (0*200, -1*200, 0*200)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void particle_step(uint32_t& tx, uint32_t& ty, uint32_t& tz,
                                              uint32_t& vx, uint32_t& vy, uint32_t& vz, float dt) {
    const float gx = __fmul_rn(0.0f, 200.0f), gy = __fmul_rn(-1.0f, 200.0f), gz = __fmul_rn(0.0f, 200.0f);
    float fvx = __uint_as_float(vx), fvy = __uint_as_float(vy), fvz = __uint_as_float(vz);
    fvx = __fadd_rn(fvx, __fmul_rn(gx, dt));   // **velocity += gravity * time_step
    fvy = __fadd_rn(fvy, __fmul_rn(gy, dt));
    fvz = __fadd_rn(fvz, __fmul_rn(gz, dt));
    float ftx = __fadd_rn(__uint_as_float(tx), __fmul_rn(fvx, dt));  // translation += **velocity * time_step
    float fty = __fadd_rn(__uint_as_float(ty), __fmul_rn(fvy, dt));
    float ftz = __fadd_rn(__uint_as_float(tz), __fmul_rn(fvz, dt));
    vx = __float_as_uint(fvx); vy = __float_as_uint(fvy); vz = __float_as_uint(fvz);
    tx = __float_as_uint(ftx); ty = __float_as_uint(fty); tz = __float_as_uint(ftz);
}

// =============================================================================================
// THE fused kernel: interprets the whole request vector (Load / Advance / Save ...) for the
// particles bundle with every entity's state in registers.  One launch per handle_requests.
//   * LOAD    : read the snapshot image once (S bytes/entity)
//   * ADVANCE : update_particles + despawn_particles in registers (0 bytes)
//   * SAVE    : stream the state into the frame's slot (S bytes/entity) and fold the per-entity
//               seahashes of the checksummed columns: warp REDUX.XOR -> shared atomics -> one
//               global atomic per block per (save, column) -> last block publishes to host memory
//   * end     : write the live image once
// Frames of one entity are sequentially dependent, so the frame loop is per thread; entities are
// independent, so the grid dimension is the entity dimension.
// =============================================================================================
template <int VEC, int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_particles_program(const __grid_constant__ ProgramParams p) {
    __shared__ unsigned long long s_acc[kMaxSaves * kAccStride];
    __shared__ unsigned int s_last;
    for (uint32_t i = threadIdx.x; i < p.n_saves * kAccStride; i += BLOCK) s_acc[i] = 0ULL;
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_groups = (p.max_rows + VEC - 1) / VEC;
    const size_t plane_bytes = size_t(p.epad) * 4u;
    const size_t alive_off = size_t(p.words) * plane_bytes;
    uint8_t* const live = p.arena;

    for (uint32_t g = blockIdx.x * BLOCK + threadIdx.x; g - lane < n_groups; g += gridDim.x * BLOCK) {
        const bool valid = g < n_groups;
        const uint32_t row0 = g * VEC;
        const size_t woff = size_t(row0) * 4u;

        // ------------------------------ active words ------------------------------
        uint32_t tr[3][VEC], vl[3][VEC], tl[2][VEC];
        uint32_t alive = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int j = 0; j < VEC; ++j) { tr[k][j] = 0; vl[k][j] = 0; }
#pragma unroll
        for (int j = 0; j < VEC; ++j) { tl[0][j] = 0; tl[1][j] = 0; }

        auto load_active = [&](const uint8_t* img, uint32_t n_rows) {
            if (valid) {
#pragma unroll
                for (int k = 0; k < 3; ++k) vec_load<VEC>(img + (p.t_plane + k) * plane_bytes + woff, tr[k]);
#pragma unroll
                for (int k = 0; k < 3; ++k) vec_load<VEC>(img + (p.v_plane + k) * plane_bytes + woff, vl[k]);
#pragma unroll
                for (int k = 0; k < 2; ++k) vec_load<VEC>(img + (p.l_plane + k) * plane_bytes + woff, tl[k]);
                alive = alive_load<VEC>(img + alive_off + row0) & rows_mask<VEC>(row0, n_rows);
            }
        };
        auto store_active = [&](uint8_t* img, uint32_t n_rows) {
            if (valid) {
#pragma unroll
                for (int k = 0; k < 3; ++k) vec_store<VEC>(img + (p.t_plane + k) * plane_bytes + woff, tr[k]);
#pragma unroll
                for (int k = 0; k < 3; ++k) vec_store<VEC>(img + (p.v_plane + k) * plane_bytes + woff, vl[k]);
#pragma unroll
                for (int k = 0; k < 2; ++k) vec_store<VEC>(img + (p.l_plane + k) * plane_bytes + woff, tl[k]);
                alive_store<VEC>(img + alive_off + row0, alive & rows_mask<VEC>(row0, n_rows));
            }
        };

        if (p.flags & PF_READ_LIVE) load_active(live, p.live_rows);

        // lane of the per-entity hash that only depends on the RollbackOrdered index
        uint64_t t0[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) t0[j] = sea_order_lane(p.order_base + row0 + j);

        for (uint32_t i = 0; i < p.n_ops; ++i) {
            const Op& op = p.ops[i];
            if (op.kind == OP_ADVANCE) {
                const float dt = __uint_as_float(op.dt_bits);
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    if ((alive >> (8 * j)) & 1u) {
                        particle_step(tr[0][j], tr[1][j], tr[2][j], vl[0][j], vl[1][j], vl[2][j], dt);
                        // despawn_particles (particles.rs:282-289): ttl -= 1 (wrapping usize); despawn at 0
                        uint32_t lo = tl[0][j], hi = tl[1][j];
                        hi -= (lo == 0u) ? 1u : 0u;
                        lo -= 1u;
                        tl[0][j] = lo; tl[1][j] = hi;
                        if ((lo | hi) == 0u) alive &= ~(0xFFu << (8 * j));
                    }
                }
            } else if (op.kind == OP_SAVE) {
                uint8_t* img = p.arena + size_t(op.image) * p.image_bytes;
                alive &= rows_mask<VEC>(row0, op.n_rows);
                if (!(op.flags & OPF_NO_STORE)) store_active(img, op.n_rows);
                // ---- checksum partials (component_checksum.rs:81-90) ----
                uint64_t hx_t = 0, hx_v = 0;
                uint32_t n_alive = 0, bad = 0;
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    if ((alive >> (8 * j)) & 1u) {
                        ++n_alive;
                        if (p.flags & PF_CK_T) {
                            if (p.flags & PF_CK_T_FINITE)
                                bad |= f32_bits_nonfinite(tr[0][j]) | f32_bits_nonfinite(tr[1][j]) | f32_bits_nonfinite(tr[2][j]);
                            uint64_t c = sea_hash_12(uint64_t(tr[0][j]) | (uint64_t(tr[1][j]) << 32), tr[2][j]);
                            hx_t ^= sea_hash_entity(t0[j], c);
                        }
                        if (p.flags & PF_CK_V) {
                            if (p.flags & PF_CK_V_FINITE)
                                bad |= f32_bits_nonfinite(vl[0][j]) | f32_bits_nonfinite(vl[1][j]) | f32_bits_nonfinite(vl[2][j]);
                            uint64_t c = sea_hash_12(uint64_t(vl[0][j]) | (uint64_t(vl[1][j]) << 32), vl[2][j]);
                            hx_v ^= sea_hash_entity(t0[j], c);
                        }
                    }
                }
                // warp-level fold (REDUX), then one shared-memory atomic per warp
                const unsigned full = 0xffffffffu;
                uint32_t tlo = __reduce_xor_sync(full, uint32_t(hx_t)), thi = __reduce_xor_sync(full, uint32_t(hx_t >> 32));
                uint32_t vlo = __reduce_xor_sync(full, uint32_t(hx_v)), vhi = __reduce_xor_sync(full, uint32_t(hx_v >> 32));
                uint32_t cnt = __reduce_add_sync(full, n_alive);
                uint32_t anybad = __reduce_or_sync(full, bad);
                if (lane == 0) {
                    unsigned long long* a = &s_acc[op.save_index * kAccStride];
                    if (p.flags & PF_CK_T) atomicXor(&a[p.ck_t_slot], (unsigned long long)tlo | ((unsigned long long)thi << 32));
                    if (p.flags & PF_CK_V) atomicXor(&a[p.ck_v_slot], (unsigned long long)vlo | ((unsigned long long)vhi << 32));
                    atomicAdd(&a[6], (unsigned long long)cnt);
                    if (anybad) atomicOr(&a[7], 1ULL);
                }
            } else {  // OP_LOAD
                const uint8_t* img = p.arena + size_t(op.image) * p.image_bytes;
                load_active(img, op.n_rows);
            }
        }
        if (p.flags & PF_WRITE_LIVE_ACTIVE) store_active(live, p.max_rows);

        // ------------------------------ passive planes ------------------------------
        // columns no compiled system writes (rotation, scale, any other registered POD): the
        // same program, but a value is only ever loaded and stored, never computed on.
        if (valid) {
            for (uint32_t pp = 0; pp < p.n_passive; pp += 4) {
                uint32_t v[4][VEC];
                const uint32_t nk = min(4u, p.n_passive - pp);
                if (p.flags & PF_READ_LIVE) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < nk) vec_load<VEC>(live + p.passive[pp + k] * plane_bytes + woff, v[k]);
                }
                for (uint32_t i = 0; i < p.n_ops; ++i) {
                    const Op& op = p.ops[i];
                    if (op.kind == OP_LOAD) {
                        const uint8_t* img = p.arena + size_t(op.image) * p.image_bytes;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (k < nk) vec_load<VEC>(img + p.passive[pp + k] * plane_bytes + woff, v[k]);
                    } else if (op.kind == OP_SAVE && !(op.flags & OPF_NO_STORE)) {
                        uint8_t* img = p.arena + size_t(op.image) * p.image_bytes;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            if (k < nk) vec_store<VEC>(img + p.passive[pp + k] * plane_bytes + woff, v[k]);
                    }
                }
                if (p.flags & PF_WRITE_LIVE_PASSIVE) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (k < nk) vec_store<VEC>(live + p.passive[pp + k] * plane_bytes + woff, v[k]);
                }
            }
        }
    }

    // ---- block partials -> global accumulators -> (last block) host-visible results ----
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < p.n_saves * kAccStride; i += BLOCK) {
        unsigned long long v = s_acc[i];
        uint32_t c = i % kAccStride;
        if (v) {
            if (c == 6) atomicAdd(&p.accum[i], v);
            else if (c == 7) atomicOr(&p.accum[i], v);
            else atomicXor(&p.accum[i], v);
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (uint32_t i = threadIdx.x; i < p.n_saves * kAccStride; i += BLOCK)
            p.out[i] = atomicExch(&p.accum[i], 0ULL);  // publish and re-arm for the next launch
        if (threadIdx.x == 0) *p.ticket = 0u;
    }
}

// =============================================================================================
// Stepwise (generic) path: one kernel per request, any registered schema / system list.
// =============================================================================================

// copy rows [0, n_rows) of every word plane + the alive plane from one image to another
// (ComponentSnapshotPlugin::save / ::load as a pure copy).  Thread = 4 consecutive rows.
__global__ void __launch_bounds__(256) k_copy_image(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                    uint32_t epad, uint32_t words, uint32_t n_rows_src,
                                                    uint32_t n_rows_copy) {
    const uint32_t n_groups = (n_rows_copy + 3) / 4;
    const size_t plane_bytes = size_t(epad) * 4u;
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += gridDim.x * blockDim.x) {
        const size_t woff = size_t(g) * 16u;
        for (uint32_t w = 0; w < words; w += 4) {
            uint4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (w + k < words) v[k] = __ldcs(reinterpret_cast<const uint4*>(src + (w + k) * plane_bytes + woff));
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (w + k < words) __stcs(reinterpret_cast<uint4*>(dst + (w + k) * plane_bytes + woff), v[k]);
        }
        uint32_t a = __ldcs(reinterpret_cast<const uint32_t*>(src + size_t(words) * plane_bytes + g * 4u));
        a &= rows_mask<4>(g * 4u, n_rows_src);
        __stcs(reinterpret_cast<uint32_t*>(dst + size_t(words) * plane_bytes + g * 4u), a);
    }
}

// per-column XOR of per-entity hashes over live rows (component_checksum.rs:67-108), generic
// byte range [off, off+len) of an element that is stored as `words` planes starting at first_plane.
// acc[col_slot] ^= ..., acc[6] += live rows (only when count_alive), acc[7] |= nonfinite.
__global__ void __launch_bounds__(256) k_checksum_column(const uint8_t* __restrict__ img, uint32_t epad,
                                                         uint32_t total_words, uint32_t first_plane,
                                                         uint32_t off, uint32_t len, uint32_t finite_flag,
                                                         uint32_t n_rows, unsigned long long order_base,
                                                         unsigned long long* acc, uint32_t col_slot,
                                                         uint32_t count_alive, uint32_t hash_column) {
    const size_t plane_bytes = size_t(epad) * 4u;
    const uint8_t* alive = img + size_t(total_words) * plane_bytes;
    const uint32_t lane = threadIdx.x & 31u;
    uint64_t hx = 0;
    uint32_t cnt = 0, bad = 0;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r - lane < n_rows; r += gridDim.x * blockDim.x) {
        if (r < n_rows && alive[r]) {
            ++cnt;
            if (hash_column) {
                auto byte_at = [&](uint32_t i) -> uint8_t {
                    uint32_t b = off + i;
                    uint32_t w = *reinterpret_cast<const uint32_t*>(img + (first_plane + (b >> 2)) * plane_bytes + size_t(r) * 4u);
                    return uint8_t(w >> (8 * (b & 3u)));
                };
                if (finite_flag)
                    for (uint32_t i = 0; i + 4 <= len; i += 4) {
                        uint32_t w = uint32_t(byte_at(i)) | (uint32_t(byte_at(i + 1)) << 8) | (uint32_t(byte_at(i + 2)) << 16) | (uint32_t(byte_at(i + 3)) << 24);
                        bad |= f32_bits_nonfinite(w);
                    }
                uint64_t custom = sea_hash_stream(len, byte_at);
                hx ^= sea_hash_2xu64(order_base + r, custom);
            }
        }
    }
    const unsigned full = 0xffffffffu;
    uint32_t lo = __reduce_xor_sync(full, uint32_t(hx)), hi = __reduce_xor_sync(full, uint32_t(hx >> 32));
    uint32_t c = __reduce_add_sync(full, cnt);
    uint32_t b = __reduce_or_sync(full, bad);
    if (lane == 0) {
        if (hash_column) atomicXor(&acc[col_slot], (unsigned long long)lo | ((unsigned long long)hi << 32));
        if (count_alive) atomicAdd(&acc[6], (unsigned long long)c);
        if (b) atomicOr(&acc[7], 1ULL);
    }
}

// copy the accumulators of n_saves saves to the host-mapped result block and re-arm them
__global__ void k_publish(unsigned long long* accum, unsigned long long* out, uint32_t n) {
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) out[i] = atomicExch(&accum[i], 0ULL);
}

// ---- GgrsSchedule systems on the live image (stepwise path) ----
__global__ void __launch_bounds__(256) k_sys_particles_update(uint8_t* img, uint32_t epad, uint32_t words,
                                                              uint32_t t_plane, uint32_t v_plane, uint32_t n_rows,
                                                              uint32_t dt_bits) {
    const size_t pb = size_t(epad) * 4u;
    const uint8_t* alive = img + size_t(words) * pb;
    const float dt = __uint_as_float(dt_bits);
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += gridDim.x * blockDim.x) {
        if (!alive[r]) continue;
        uint32_t* t[3]; uint32_t* v[3];
        for (int k = 0; k < 3; ++k) {
            t[k] = reinterpret_cast<uint32_t*>(img + (t_plane + k) * pb) + r;
            v[k] = reinterpret_cast<uint32_t*>(img + (v_plane + k) * pb) + r;
        }
        uint32_t tx = *t[0], ty = *t[1], tz = *t[2], vx = *v[0], vy = *v[1], vz = *v[2];
        particle_step(tx, ty, tz, vx, vy, vz, dt);
        *t[0] = tx; *t[1] = ty; *t[2] = tz; *v[0] = vx; *v[1] = vy; *v[2] = vz;
    }
}

// `alive_next` receives the despawns; it is applied after every system of the schedule has run
// (Commands are deferred to the end of GgrsSchedule).
__global__ void __launch_bounds__(256) k_sys_particles_despawn(uint8_t* img, uint32_t epad, uint32_t words,
                                                               uint32_t l_plane, uint32_t n_rows, uint8_t* kill) {
    const size_t pb = size_t(epad) * 4u;
    const uint8_t* alive = img + size_t(words) * pb;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += gridDim.x * blockDim.x) {
        if (!alive[r]) continue;
        uint32_t* lo = reinterpret_cast<uint32_t*>(img + l_plane * pb) + r;
        uint32_t* hi = reinterpret_cast<uint32_t*>(img + (l_plane + 1) * pb) + r;
        uint64_t ttl = (uint64_t(*hi) << 32) | *lo;
        ttl -= 1;
        *lo = uint32_t(ttl); *hi = uint32_t(ttl >> 32);
        if (ttl == 0) kill[r] = 1;
    }
}

// x.0 += k   (tests/component_rollback.rs:25-29)
__global__ void __launch_bounds__(256) k_sys_u32_add(uint8_t* img, uint32_t epad, uint32_t words, uint32_t plane,
                                                     uint32_t n_rows, uint32_t k) {
    const size_t pb = size_t(epad) * 4u;
    const uint8_t* alive = img + size_t(words) * pb;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += gridDim.x * blockDim.x)
        if (alive[r]) reinterpret_cast<uint32_t*>(img + plane * pb)[r] += k;
}

// h = h.saturating_sub(k); despawn at 0   (tests/synctest.rs:38-45)
__global__ void __launch_bounds__(256) k_sys_u32_satsub_despawn(uint8_t* img, uint32_t epad, uint32_t words,
                                                                uint32_t plane, uint32_t n_rows, uint32_t k, uint8_t* kill) {
    const size_t pb = size_t(epad) * 4u;
    const uint8_t* alive = img + size_t(words) * pb;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += gridDim.x * blockDim.x) {
        if (!alive[r]) continue;
        uint32_t* x = reinterpret_cast<uint32_t*>(img + plane * pb) + r;
        uint32_t v = *x;
        v = v > k ? v - k : 0u;
        *x = v;
        if (v == 0) kill[r] = 1;
    }
}

// c.0 = count  (the deliberately non-deterministic system of tests/synctest.rs:92-97)
__global__ void __launch_bounds__(256) k_sys_u32_store(uint8_t* img, uint32_t epad, uint32_t words, uint32_t plane,
                                                       uint32_t n_rows, uint32_t value) {
    const size_t pb = size_t(epad) * 4u;
    const uint8_t* alive = img + size_t(words) * pb;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += gridDim.x * blockDim.x)
        if (alive[r]) reinterpret_cast<uint32_t*>(img + plane * pb)[r] = value;
}

// apply deferred despawn commands: alive &= !kill ; kill = 0
__global__ void __launch_bounds__(256) k_apply_despawns(uint8_t* img, uint32_t epad, uint32_t words, uint32_t n_rows,
                                                        uint8_t* kill) {
    uint8_t* alive = img + size_t(words) * size_t(epad) * 4u;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += gridDim.x * blockDim.x)
        if (kill[r]) { alive[r] = 0; kill[r] = 0; }
}

}  // namespace bgr
