// Snapshot save / restore as a TMA-staged bulk copy (sm_90+/sm_100a `cp.async.bulk`, SASS UBLKCP):
//
//     HBM image A --cp.async.bulk(global->shared, mbarrier complete_tx)--> shared-memory stage
//                 --cp.async.bulk(shared->global, bulk_group)-----------> HBM image B
//
// No register staging: one elected thread moves a whole tile of every word plane (and the alive
// plane) with a handful of bulk-copy instructions, a ring of stages keeps several tiles in flight
// per SM, and — for a Save — all threads hash the checksummed byte ranges straight out of the staged
// tile (planar layout => conflict-free shared loads) while the TMA store streams it to the slot.
//
// Reference semantics: ComponentSnapshotPlugin::save (component_snapshot.rs:66-84) + the checksum
// systems of SaveWorld (component_checksum.rs:67-108, entity_checksum.rs:29-52) in ONE launch;
// ComponentSnapshotPlugin::load (component_snapshot.rs:95-123) as the reverse copy.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "kernels.cuh"
#include "seahash.cuh"

namespace bgr {

constexpr int kTmaStages = 3;
constexpr int kTmaBlock = 256;
constexpr int kMaxHashCols = 6;

struct HashSpec {
    uint32_t first_plane, off, len, finite, slot;
};

struct TmaCopyParams {
    const uint8_t* src;
    uint8_t* dst;
    unsigned long long order_base;
    unsigned long long* accum;  // device row of this save (kAccStride u64) or nullptr
    unsigned long long* out;    // host-mapped row to publish to (or nullptr: publish later)
    unsigned int* ticket;
    uint32_t epad, words, tile_rows;
    uint32_t n_rows_src;        // rows that exist in the source image (alive beyond is forced 0)
    uint32_t n_rows_copy;       // rows to move (>= n_rows_src when a Load shrinks the world)
    uint32_t n_hash;            // 0 for Load
    uint32_t count_alive;       // Save: fold live-row count into accum[6]
    uint32_t store;             // 0: checksum only (ring depth 0)
    HashSpec hash[kMaxHashCols];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// stage layout: [plane 0: tile_rows*4 B] ... [plane W-1] [alive: tile_rows B]
__global__ void __launch_bounds__(kTmaBlock, 1) k_image_tma(const __grid_constant__ TmaCopyParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[kTmaStages];
    __shared__ unsigned int s_last;

    const uint32_t T = p.tile_rows;
    const size_t stage_bytes = size_t(T) * (size_t(p.words) * 4u + 1u);
    const size_t plane_bytes = size_t(p.epad) * 4u;
    const size_t alive_off = size_t(p.words) * plane_bytes;
    const uint32_t n_tiles = (p.n_rows_copy + T - 1) / T;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;

    if (tid == 0) {
        for (int s = 0; s < kTmaStages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto tile_rows_of = [&](uint32_t tile) { return min(T, p.n_rows_copy - tile * T); };
    auto issue_load = [&](uint32_t tile, uint32_t stage) {  // thread 0 only
        const uint32_t rows = tile_rows_of(tile);
        const uint32_t wbytes = ((rows * 4u) + 15u) & ~15u;
        const uint32_t abytes = (rows + 15u) & ~15u;
        uint8_t* st = smem + stage * stage_bytes;
        mbar_arrive_expect_tx(&full[stage], wbytes * p.words + abytes);
        const size_t row_off = size_t(tile) * T;
        for (uint32_t w = 0; w < p.words; ++w)
            tma_load_1d(st + size_t(w) * T * 4u, p.src + w * plane_bytes + row_off * 4u, wbytes, &full[stage]);
        tma_load_1d(st + size_t(p.words) * T * 4u, p.src + alive_off + row_off, abytes, &full[stage]);
    };

    // my tiles: blockIdx.x, blockIdx.x + gridDim.x, ...
    const uint32_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (tid == 0)
        for (uint32_t k = 0; k < my_tiles && k < kTmaStages; ++k) issue_load(blockIdx.x + k * gridDim.x, k);

    uint64_t hx[kMaxHashCols];
#pragma unroll
    for (int c = 0; c < kMaxHashCols; ++c) hx[c] = 0;
    uint32_t n_alive = 0, bad = 0;

    for (uint32_t k = 0; k < my_tiles; ++k) {
        const uint32_t stage = k % kTmaStages, parity = (k / kTmaStages) & 1u;
        const uint32_t tile = blockIdx.x + k * gridDim.x;
        const uint32_t rows = tile_rows_of(tile);
        const uint32_t row_base = tile * T;
        uint8_t* st = smem + stage * stage_bytes;
        uint8_t* st_alive = st + size_t(p.words) * T * 4u;
        mbar_wait(&full[stage], parity);

        if (row_base + rows > p.n_rows_src) {
            // a Load that shrinks the world: rows the snapshot never contained must come back dead
            for (uint32_t r = tid; r < ((rows + 15u) & ~15u); r += kTmaBlock)
                if (row_base + r >= p.n_rows_src) st_alive[r] = 0;
            fence_proxy_async();
            __syncthreads();
        }
        if (tid == 0 && p.store) {
            const uint32_t wbytes = ((rows * 4u) + 15u) & ~15u;
            const uint32_t abytes = (rows + 15u) & ~15u;
            for (uint32_t w = 0; w < p.words; ++w)
                tma_store_1d(p.dst + w * plane_bytes + size_t(row_base) * 4u, st + size_t(w) * T * 4u, wbytes);
            tma_store_1d(p.dst + alive_off + row_base, st_alive, abytes);
            tma_commit();
        }
        if (p.n_hash || p.count_alive) {
            for (uint32_t r = tid; r < rows; r += kTmaBlock) {
                if (!st_alive[r]) continue;
                ++n_alive;
                const uint64_t t0 = sea_order_lane(p.order_base + row_base + r);
#pragma unroll
                for (int c = 0; c < kMaxHashCols; ++c) {
                    if (c < p.n_hash) {
                        const HashSpec hs = p.hash[c];
                        const uint32_t* col = reinterpret_cast<const uint32_t*>(st) + size_t(hs.first_plane) * T + r;
                        uint64_t custom;
                        if (hs.off == 0 && hs.len == 12) {  // 3 x u32 `to_bits` fields (particles.rs:107-120, 207-222)
                            uint32_t a = col[0], b = col[T], d = col[2 * T];
                            if (hs.finite) bad |= f32_bits_nonfinite(a) | f32_bits_nonfinite(b) | f32_bits_nonfinite(d);
                            custom = sea_hash_12(uint64_t(a) | (uint64_t(b) << 32), d);
                        } else {
                            auto byte_at = [&](uint32_t i) -> uint8_t {
                                uint32_t bb = hs.off + i;
                                return uint8_t(col[size_t(bb >> 2) * T] >> (8 * (bb & 3u)));
                            };
                            if (hs.finite)
                                for (uint32_t i = 0; i + 4 <= hs.len; i += 4) bad |= f32_bits_nonfinite(col[size_t((hs.off + i) >> 2) * T]);
                            custom = sea_hash_stream(hs.len, byte_at);
                        }
                        hx[c] ^= sea_hash_entity(t0, custom);
                    }
                }
            }
        }
        __syncthreads();  // everyone is done reading this stage
        if (tid == 0 && k + kTmaStages < my_tiles) {
            tma_wait_read_all();  // the bulk store has finished reading the stage
            issue_load(blockIdx.x + (k + kTmaStages) * gridDim.x, stage);
        }
    }
    if (tid == 0) tma_wait_all();  // all stores have landed before the kernel (and its results) complete

    if (p.accum) {
        const unsigned full_mask = 0xffffffffu;
#pragma unroll
        for (int c = 0; c < kMaxHashCols; ++c) {
            if (c < p.n_hash) {
                uint32_t lo = __reduce_xor_sync(full_mask, uint32_t(hx[c])), hi = __reduce_xor_sync(full_mask, uint32_t(hx[c] >> 32));
                if (lane == 0 && (lo | hi)) atomicXor(&p.accum[p.hash[c].slot], (unsigned long long)lo | ((unsigned long long)hi << 32));
            }
        }
        uint32_t cnt = __reduce_add_sync(full_mask, n_alive);
        uint32_t anybad = __reduce_or_sync(full_mask, bad);
        if (lane == 0) {
            if (p.count_alive && cnt) atomicAdd(&p.accum[6], (unsigned long long)cnt);
            if (anybad) atomicOr(&p.accum[7], 1ULL);
        }
        if (p.out) {
            __threadfence();
            __syncthreads();
            if (tid == 0) s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1u);
            __syncthreads();
            if (s_last) {
                __threadfence();
                for (uint32_t i = tid; i < kAccStride; i += kTmaBlock) p.out[i] = atomicExch(&p.accum[i], 0ULL);
                if (tid == 0) *p.ticket = 0u;
            }
        }
    }
}

}  // namespace bgr
