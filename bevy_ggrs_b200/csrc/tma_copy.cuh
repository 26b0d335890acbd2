// Snapshot save / restore as a TMA-staged bulk copy (sm_90+/sm_100a `cp.async.bulk`, SASS UBLKCP):
//
//     HBM image A --cp.async.bulk(global->shared, mbarrier complete_tx)--> shared-memory stage
//                 --cp.async.bulk(shared->global, bulk_group)-----------> HBM image B
//
// Images are tile-planar (kernels.cuh), so a group of tiles is ONE contiguous chunk: one elected
// thread moves ~60 KB with a single bulk-copy instruction each way, a ring of stages keeps several
// chunks in flight per SM, and no byte of the copy ever passes through a register.  For a Save all
// threads hash the checksummed byte ranges straight out of the staged tiles (planar layout =>
// conflict-free shared loads) while the TMA store streams them to the slot.
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
    uint32_t absent;  // absent bit of an optional column (0: the column is on every live row)
};

struct TmaCopyParams {
    const uint8_t* src;
    uint8_t* dst;
    unsigned long long order_base;
    unsigned long long* accum;  // device row of this save (kAccStride u64) or nullptr
    uint32_t words, tile_bytes, stage_tiles;
    uint32_t n_tiles;           // tiles to move
    uint32_t n_rows_src;        // rows that exist in the source image (alive beyond is forced 0)
    uint32_t n_hash;            // 0 for Load
    uint32_t count_alive;       // Save: fold live-row count into accum[6]
    uint32_t store;             // 0: checksum only (ring depth 0)
    HashSpec hash[kMaxHashCols];
};

__global__ void __launch_bounds__(kTmaBlock, 1) k_image_tma(const __grid_constant__ TmaCopyParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[kTmaStages];

    const uint32_t G = p.stage_tiles;
    const size_t stage_bytes = size_t(G) * p.tile_bytes;
    const uint32_t n_chunks = (p.n_tiles + G - 1) / G;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;

    if (tid == 0) {
        for (int s = 0; s < kTmaStages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto chunk_tiles = [&](uint32_t chunk) { return min(G, p.n_tiles - chunk * G); };
    auto issue_load = [&](uint32_t chunk, uint32_t stage) {  // thread 0 only
        const uint32_t bytes = chunk_tiles(chunk) * p.tile_bytes;
        mbar_arrive_expect_tx(&full[stage], bytes);
        tma_load_1d(smem + stage * stage_bytes, p.src + size_t(chunk) * stage_bytes, bytes, &full[stage]);
    };

    const uint32_t my_chunks = (n_chunks > blockIdx.x) ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (tid == 0)
        for (uint32_t k = 0; k < my_chunks && k < kTmaStages; ++k) issue_load(blockIdx.x + k * gridDim.x, k);

    uint64_t hx[kMaxHashCols];
#pragma unroll
    for (int c = 0; c < kMaxHashCols; ++c) hx[c] = 0;
    uint32_t n_alive = 0, bad = 0;

    for (uint32_t k = 0; k < my_chunks; ++k) {
        const uint32_t stage = k % kTmaStages, parity = (k / kTmaStages) & 1u;
        const uint32_t chunk = blockIdx.x + k * gridDim.x;
        const uint32_t tiles = chunk_tiles(chunk);
        const uint32_t rows = tiles * kTileRows;
        const uint32_t row_base = chunk * G * kTileRows;
        uint8_t* st = smem + stage * stage_bytes;
        mbar_wait(&full[stage], parity);

        if (row_base + rows > p.n_rows_src) {
            // a Load that shrinks the world: rows the snapshot never contained must come back dead
            for (uint32_t r = tid; r < rows; r += kTmaBlock)
                if (row_base + r >= p.n_rows_src) st[size_t(r / kTileRows) * p.tile_bytes + size_t(p.words) * kPlaneBytes + (r % kTileRows)] = 0;
            fence_proxy_async();
            __syncthreads();
        }
        if (tid == 0 && p.store) {
            tma_store_1d(p.dst + size_t(chunk) * stage_bytes, st, tiles * p.tile_bytes);
            tma_commit();
        }
        if (p.n_hash || p.count_alive) {
            for (uint32_t r = tid; r < rows; r += kTmaBlock) {
                const uint8_t* tile = st + size_t(r / kTileRows) * p.tile_bytes;
                const uint32_t i = r % kTileRows;
                const uint32_t m = tile[size_t(p.words) * kPlaneBytes + i];
                if (!(m & 1u)) continue;
                ++n_alive;
                const uint64_t t0 = sea_order_lane(p.order_base + row_base + r);
#pragma unroll
                for (int c = 0; c < kMaxHashCols; ++c) {
                    if (c < p.n_hash) {
                        const HashSpec hs = p.hash[c];
                        if (m & hs.absent) continue;  // Query<(&RollbackId, &T)> does not match this entity
                        const uint32_t* col = reinterpret_cast<const uint32_t*>(tile + size_t(hs.first_plane) * kPlaneBytes) + i;
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
        }
        __syncthreads();  // everyone is done reading this stage
        if (tid == 0 && k + kTmaStages < my_chunks) {
            tma_wait_read<0>();  // the bulk store has finished reading the stage
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
    }
}

}  // namespace bgr
