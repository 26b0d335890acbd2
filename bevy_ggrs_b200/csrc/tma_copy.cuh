// Snapshot save / restore as a TMA-staged bulk copy (sm_90+/sm_100a `cp.async.bulk`, SASS UBLKCP):
//
//     HBM image A --cp.async.bulk(global->shared, mbarrier complete_tx)--> shared-memory stage
//                 --cp.async.bulk(shared->global, bulk_group)-----------> HBM image B
//
// Images are tile-planar (kernels.cuh), so a tile is ONE contiguous chunk (31 KB for the stress schema): a dedicated
// producer thread moves it with a single bulk-copy instruction each way, a ring of up to six stages keeps five loads
// and a store in flight per SM, and no byte of the copy ever passes through a register.  For a Save sixteen consumer
// warps hash the checksummed byte ranges straight out of the staged tile (planar layout => conflict-free shared
// loads) while the TMA store streams it to the slot; producer and consumers are decoupled by full / hashed mbarriers
// per stage, tiles are claimed from a global counter (no tail imbalance between SMs).
//
// Round-1 version and what the profile said (profiles/r01_k_image_tma.metrics.csv): one block of
// 256 threads per SM did everything; thread 0 waited for the store of a stage (`wait_group.read 0`) right after issuing
// it, and 8 warps per SM could not issue the Save's hash instructions fast enough (sm__warps_active 12.5 %): 35 us per
// 1M-entity Save against 25 us for a Load.
//
// Reference semantics: ComponentSnapshotPlugin::save (component_snapshot.rs:66-84) + the checksum
// systems of SaveWorld (component_checksum.rs:67-108, entity_checksum.rs:29-52) in ONE launch;
// ComponentSnapshotPlugin::load (component_snapshot.rs:95-123) as the reverse copy.
#pragma once
#ifdef __CUDACC_RTC__  // NVRTC (the engine's run-time specialisation, generic_program_jit.cuh): no host headers
#include "rtc_prelude.cuh"
#else
#include <cuda_runtime.h>
#include <cstdint>
#endif

#include "kernels.cuh"
#include "seahash.cuh"

namespace bgr {

constexpr int kTmaMaxStages = 6;
constexpr int kTmaConsumers = kTileRows;          // one row of the staged tile per consumer thread
constexpr int kTmaBlock = kTmaConsumers + 32;     // + the producer warp
constexpr int kMaxHashCols = 6;
constexpr uint32_t kTmaEnd = 0xffffffffu;

struct HashSpec {
    uint32_t first_plane, off, len, finite, slot;
    uint32_t absent;  // absent bit of an optional column (0: the column is on every live row)
};

struct TmaCopyParams {
    const uint8_t* src;
    uint8_t* dst;
    unsigned long long order_base;
    unsigned long long* accum;  // device row of this save (kAccStride u64) or nullptr
    unsigned int* ticket;       // [0] finished blocks, [1] next tile to claim (re-armed by the last block)
    uint32_t words, tile_bytes, stages;
    uint32_t n_tiles;           // tiles to move
    uint32_t n_rows_src;        // rows that exist in the source image (alive beyond is forced 0)
    uint32_t n_hash;            // 0 for Load
    uint32_t count_alive;       // Save: fold live-row count into accum[6]
    uint32_t store;             // 0: checksum only (ring depth 0)
    HashSpec hash[kMaxHashCols];
};

#ifndef __CUDACC_RTC__
__global__ void __launch_bounds__(kTmaBlock, 1) k_image_tma(const __grid_constant__ TmaCopyParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[kTmaMaxStages], hashed[kTmaMaxStages];
    __shared__ uint32_t s_tile[kTmaMaxStages];
    __shared__ unsigned int s_acc[2 * kMaxHashCols + 2];
    __shared__ unsigned int s_last;

    const uint32_t S = p.stages;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // consumers have work when the launch hashes / counts, or when a Load shrinks the world (alive bytes of rows the
    // snapshot never contained must come back dead before the tile is stored)
    const bool consume = p.n_hash || p.count_alive || size_t(p.n_tiles) * kTileRows > p.n_rows_src;

    if (tid == 0) {
        for (uint32_t s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&hashed[s], kTmaConsumers / 32); }
        fence_mbar_init();
    }
    if (tid < 2 * kMaxHashCols + 2) s_acc[tid] = 0u;
    __syncthreads();

    auto needs_fix = [&](uint32_t tile) { return size_t(tile + 1) * kTileRows > p.n_rows_src; };

    if (warp == kTmaConsumers / 32) {
        // ------------------------------ producer: one thread drives every bulk copy ------------------------------
        if (lane == 0) {
            bool ended = false;
            uint32_t loaded = 0;  // iterations whose load (or end marker) has been issued; iteration i uses stage i % S
            // Tiles are claimed TWO issues ahead of their use: the first one is the block's own index, later ones come
            // from the global counter, and the atomic's round trip (~1 us, once per 31 KB tile — it cost a third of a
            // 1M-entity Load when it sat on the producer's critical path) is hidden behind the previous tile's copies.
            uint32_t claim = blockIdx.x;
            uint32_t pending = gridDim.x + atomicAdd(&p.ticket[1], 1u);
            auto issue = [&]() {
                const uint32_t s = loaded % S;
                const uint32_t tile = claim;
                claim = pending;
                if (tile < p.n_tiles) pending = gridDim.x + atomicAdd(&p.ticket[1], 1u);
                if (tile >= p.n_tiles) {
                    s_tile[s] = kTmaEnd;
                    mbar_arrive(&full[s]);
                    ended = true;
                } else {
                    s_tile[s] = tile;
                    mbar_arrive_expect_tx(&full[s], p.tile_bytes);
                    tma_load_1d(smem + size_t(s) * p.tile_bytes, p.src + size_t(tile) * p.tile_bytes, p.tile_bytes, &full[s]);
                }
                ++loaded;
            };
            while (loaded < S && !ended) issue();
            for (uint32_t i = 0;; ++i) {
                const uint32_t s = i % S, par = (i / S) & 1u;
                mbar_wait(&full[s], par);
                const uint32_t tile = s_tile[s];
                if (tile == kTmaEnd) break;
                if (p.store) {
                    if (consume && needs_fix(tile)) mbar_wait(&hashed[s], par);  // alive bytes were patched in the stage
                    tma_store_1d(p.dst + size_t(tile) * p.tile_bytes, smem + size_t(s) * p.tile_bytes, p.tile_bytes);
                    tma_commit();
                }
                if (i >= 1 && !ended) {  // refill the stage of iteration i-1: its store has been read, its rows hashed
                    const uint32_t sp = (i - 1) % S, parp = ((i - 1) / S) & 1u;
                    if (p.store) tma_wait_read<1>();  // only the store issued just above may still be reading
                    if (consume) mbar_wait(&hashed[sp], parp);
                    issue();
                }
            }
            tma_wait_all();  // every store has landed before the kernel (and its results) complete
        }
    } else if (consume) {
        // ------------------------------ consumers: one row of the staged tile per thread ------------------------------
        uint64_t hx[kMaxHashCols];
#pragma unroll
        for (int c = 0; c < kMaxHashCols; ++c) hx[c] = 0;
        uint32_t n_alive = 0, bad = 0;
        for (uint32_t i = 0;; ++i) {
            const uint32_t s = i % S, par = (i / S) & 1u;
            mbar_wait(&full[s], par);
            const uint32_t tile_idx = s_tile[s];
            if (tile_idx == kTmaEnd) break;
            uint8_t* tile = smem + size_t(s) * p.tile_bytes;
            const uint32_t row = tile_idx * kTileRows + tid;
            uint32_t m = tile[size_t(p.words) * kPlaneBytes + tid];
            if (needs_fix(tile_idx)) {
                if (row >= p.n_rows_src) { m = 0; tile[size_t(p.words) * kPlaneBytes + tid] = 0; }
                fence_proxy_async();  // the patched bytes are visible to the bulk store that follows
            }
            if ((p.n_hash || p.count_alive) && (m & 1u)) {
                ++n_alive;
                const uint64_t t0 = sea_order_lane(p.order_base + row);
#pragma unroll
                for (int c = 0; c < kMaxHashCols; ++c) {
                    if (c < p.n_hash) {
                        const HashSpec hs = p.hash[c];
                        if (m & hs.absent) continue;  // Query<(&RollbackId, &T)> does not match this entity
                        const uint32_t* col = reinterpret_cast<const uint32_t*>(tile + size_t(hs.first_plane) * kPlaneBytes) + tid;
                        uint64_t custom;
                        if (hs.off == 0 && hs.len == 12) {  // 3 x u32 `to_bits` fields (particles.rs:107-120, 207-222)
                            uint32_t a = col[0], b = col[kTileRows], d = col[2 * kTileRows];
                            if (hs.finite) bad |= f32_bits_nonfinite(a) | f32_bits_nonfinite(b) | f32_bits_nonfinite(d);
                            custom = sea_hash_12(uint64_t(a) | (uint64_t(b) << 32), d);
                        } else {
                            auto byte_at = [&](uint32_t q) -> uint8_t {
                                uint32_t bb = hs.off + q;
                                return uint8_t(col[size_t(bb >> 2) * kTileRows] >> (8 * (bb & 3u)));
                            };
                            if (hs.finite)
                                for (uint32_t q = 0; q + 4 <= hs.len; q += 4) bad |= f32_bits_nonfinite(col[size_t((hs.off + q) >> 2) * kTileRows]);
                            custom = sea_hash_stream(hs.len, byte_at);
                        }
                        hx[c] ^= sea_hash_entity(t0, custom);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&hashed[s]);  // this warp is done with the stage
        }
        if (p.accum) {
            const unsigned full_mask = 0xffffffffu;
#pragma unroll
            for (int c = 0; c < kMaxHashCols; ++c) {
                if (c < p.n_hash) {
                    uint32_t lo = __reduce_xor_sync(full_mask, uint32_t(hx[c])), hi = __reduce_xor_sync(full_mask, uint32_t(hx[c] >> 32));
                    if (lane == 0) { if (lo) atomicXor(&s_acc[2 * c], lo); if (hi) atomicXor(&s_acc[2 * c + 1], hi); }
                }
            }
            uint32_t cnt = __reduce_add_sync(full_mask, n_alive);
            uint32_t anybad = __reduce_or_sync(full_mask, bad);
            if (lane == 0) {
                if (cnt) atomicAdd(&s_acc[2 * kMaxHashCols], cnt);
                if (anybad) atomicOr(&s_acc[2 * kMaxHashCols + 1], 1u);
            }
        }
    }
    __syncthreads();
    if (p.accum && tid < kMaxHashCols + 2) {
        if (tid < kMaxHashCols) {
            const unsigned long long v = (unsigned long long)s_acc[2 * tid] | ((unsigned long long)s_acc[2 * tid + 1] << 32);
            if (tid < p.n_hash && v) atomicXor(&p.accum[p.hash[tid].slot], v);
        } else if (tid == kMaxHashCols) {
            if (p.count_alive && s_acc[2 * kMaxHashCols]) atomicAdd(&p.accum[6], (unsigned long long)s_acc[2 * kMaxHashCols]);
        } else if (s_acc[2 * kMaxHashCols + 1]) {
            atomicOr(&p.accum[7], 1ULL);
        }
    }
    // re-arm the tile counter for the next launch on this stream
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&p.ticket[0], 1u) == gridDim.x - 1u);
    __syncthreads();
    if (s_last && tid == 0) { p.ticket[0] = 0u; p.ticket[1] = 0u; }
}

#endif  // !__CUDACC_RTC__

}  // namespace bgr
