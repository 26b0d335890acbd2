// The one-launch path for ANY registered schema and the compiled GgrsSchedule systems: a whole Vec<GgrsRequest>
// (Load / Advance / Save ...) interpreted by one kernel, like k_particles_program, but without knowing the columns at
// compile time.  One 512-row tile per block iteration lives in SHARED MEMORY for the whole program:
//
//     cp.async.bulk  slot/live image --> shared tile              (LOAD, or the program's first read)
//     ADVANCE : every registered system updates its rows of the shared tile in place (a thread owns its rows for the
//               whole program, so no barrier separates systems; despawn commands are applied after the last system)
//     SAVE    : cp.async.bulk shared tile --> the frame's slot, while all threads hash the checksummed byte ranges
//               of their rows out of the same tile (component_checksum.rs:67-108) and count live rows
//     end     : cp.async.bulk shared tile --> live image
//
// so a SyncTest tick of a box_game / score-and-health style world is ONE launch instead of one launch per request
// and per system (round 1's "stepwise" path, which stays as the fallback for schemas wider than a shared-memory tile).
// Per-entity component presence (BGR_STRATEGY_OPTIONAL) is the row's mask byte: it travels with the tile, and every
// system / checksum applies the reference's query filter per row (kernels.cuh row_matches).
//
// Reference semantics: handle_requests (schedule_systems.rs:170-289) over ComponentSnapshotPlugin::save / load
// (component_snapshot.rs:66-123), the checksum plugins, and the systems listed in include/bevy_ggrs_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/bevy_ggrs_b200.h"
#include "kernels.cuh"
#include "seahash.cuh"
#include "tma_copy.cuh"

namespace bgr {

constexpr int kMaxGenericSys = 8;
// threads per block: 256 (two rows per thread) for worlds with many tiles per SM, 512 (one row per thread, twice the
// warps per tile) for small worlds, where a tick is one wave of blocks and latency-bound per warp (ncu, 100k entities:
// 2.6 warps per scheduler, issue slots 24 % busy)

struct SysSpec {
    uint32_t id;      // bgr_system
    uint32_t plane0;  // first word plane the system touches (column's first plane + byte_offset / 4)
    uint32_t plane1;  // second bound column's first plane (Velocity for the Transform/Velocity systems)
    uint32_t need;    // absent bits of the bound columns: the query matches a row iff row_matches(mask, need)
    uint32_t param;   // k (U32_ADD / U32_SATSUB_DESPAWN) or the system's index among the call-count systems
};

struct GenericParams {
    uint8_t* arena;
    unsigned long long order_base;
    unsigned long long* accum;  // device [kMaxSaves][kAccStride]
    unsigned long long* out;    // host-mapped result block (same layout / protocol as k_particles_program)
    unsigned int* ticket;       // [0] block-completion ticket, [1] dynamic tile counter
    unsigned long long seq;
    unsigned long long* trace;
    uint32_t words, tile_bytes, n_ops, n_saves, n_tiles, live_rows, flags, n_hash, n_sys;
    HashSpec hash[kMaxHashCols];
    SysSpec sys[kMaxGenericSys];
    Op ops[kMaxOps];
};
static_assert(sizeof(GenericParams) <= 4000, "kernel parameter block must fit 4 KB");

// seahash of bytes [off, off+len) of one row's element whose words are `col[w * kTileRows]` (a column of the shared tile)
// (__noinline__: inlined per checksummed column and per row the interpreter grew to 11k instructions — 176 KB of code,
// more than the SM's instruction cache — and ran 2.7x slower per frame than the specialised bundle kernel)
__device__ __noinline__ uint64_t hash_row_range(const uint32_t* col, uint32_t off, uint32_t len) {
    if (((off | len) & 3u) == 0u) {  // word-aligned range (every POD of u32 / f32 / u64 fields): no byte shuffling
        const uint32_t* w = col + size_t(off >> 2) * kTileRows;
        // the common element sizes without a loop (warp-uniform switch: every row of a column has the same range)
        switch (len) {
        case 4: return sea_diffuse(sea_diffuse(kSeaA ^ uint64_t(w[0])) ^ kSeaB ^ kSeaC ^ kSeaD ^ 4ULL);
        case 8: return sea_hash_u64(uint64_t(w[0]) | (uint64_t(w[kTileRows]) << 32));
        case 12: return sea_hash_12(uint64_t(w[0]) | (uint64_t(w[kTileRows]) << 32), w[2 * kTileRows]);
        case 16: return sea_hash_2xu64(uint64_t(w[0]) | (uint64_t(w[kTileRows]) << 32), uint64_t(w[2 * kTileRows]) | (uint64_t(w[3 * kTileRows]) << 32));
        default: break;
        }
        uint64_t a = kSeaA, b = kSeaB, c = kSeaC, d = kSeaD;
        uint32_t i = 0;
        for (; i + 8 <= len; i += 8) {
            const uint64_t x = uint64_t(w[0]) | (uint64_t(w[kTileRows]) << 32);
            w += 2 * kTileRows;
            const uint64_t t = sea_diffuse(a ^ x);
            a = b; b = c; c = d; d = t;
        }
        if (i < len) a = sea_diffuse(a ^ uint64_t(w[0]));
        return sea_diffuse(a ^ b ^ c ^ d ^ uint64_t(len));
    }
    auto byte_at = [&](uint32_t q) -> uint8_t {
        const uint32_t bb = off + q;
        return uint8_t(col[size_t(bb >> 2) * kTileRows] >> (8 * (bb & 3u)));
    };
    return sea_hash_stream(len, byte_at);
}

template <int kGenericBlock>
__global__ void __launch_bounds__(kGenericBlock) k_generic_program(const __grid_constant__ GenericParams p) {
    constexpr int kGenericRowsPerThread = kTileRows / kGenericBlock;
    extern __shared__ __align__(128) uint8_t s_buf[];  // TWO tile buffers (ping-pong), each tile_bytes rounded up to 128
    __shared__ unsigned int s_acc[kMaxSaves * kAccStride * 2];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_next;
    __shared__ unsigned int s_last;

    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    if (p.trace && tid == 0) atomicMin(&p.trace[0], globaltimer_ns());
    for (uint32_t i = tid; i < p.n_saves * kAccStride * 2; i += kGenericBlock) s_acc[i] = 0u;
    if (tid == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); }
    __syncthreads();

    // Ping-pong: an ADVANCE reads the current buffer and writes the other one, so the bulk store a SAVE issued from the
    // current buffer keeps reading it while the next frame is already being computed — with ONE buffer every frame waited
    // for its predecessor's store to finish reading shared memory (measured: 4.5 us per frame instead of ~2).
    const uint32_t buf_stride = (p.tile_bytes + 127u) & ~127u;
    uint32_t cur = 0;                       // buffer holding the tile's current state (block-uniform)
    uint32_t phase = 0;
    uint32_t n_stores = 0;                  // bulk stores committed so far by thread 0 (block-uniform count)
    uint32_t last_store[2] = {0, 0};        // n_stores right after the last store that reads each buffer (0: none)
    auto tile_ptr = [&](uint32_t b) { return s_buf + size_t(b) * buf_stride; };
    auto alive_ptr = [&](uint32_t b) { return tile_ptr(b) + size_t(p.words) * kPlaneBytes; };

    // buffer b is about to be overwritten: every bulk store that reads it must be done reading, and every thread must be
    // done with its previous content
    auto before_write = [&](uint32_t b) {
        if (last_store[b] != 0) {
            if (tid == 0) {
                if (n_stores - last_store[b] == 0) tma_wait_read<0>();   // the most recent store reads b
                else tma_wait_read<1>();                                  // at least one newer store exists: all but the newest are done
            }
            last_store[b] = 0;
        }
        __syncthreads();
    };
    // bring tile `t` of image `img` into the OTHER buffer and make it current; rows the image never contained come back dead
    auto load_tile = [&](const uint8_t* img, uint32_t t, uint32_t n_rows_src) {
        const uint32_t nb = cur ^ 1u;
        before_write(nb);
        if (tid == 0) {
            mbar_arrive_expect_tx(&s_bar, p.tile_bytes);
            tma_load_1d(tile_ptr(nb), img + size_t(t) * p.tile_bytes, p.tile_bytes, &s_bar);
        }
        mbar_wait(&s_bar, phase);
        phase ^= 1u;
        cur = nb;
        if (size_t(t + 1) * kTileRows > n_rows_src) {
            uint8_t* al = alive_ptr(cur);
#pragma unroll
            for (int k = 0; k < kGenericRowsPerThread; ++k) {
                const uint32_t r = tid + k * kGenericBlock;
                if (t * kTileRows + r >= n_rows_src) al[r] = 0;
            }
        }
    };
    auto store_tile = [&](uint8_t* img, uint32_t t) {
        fence_proxy_async();  // generic-proxy writes of this thread are visible to the bulk (async-proxy) store
        __syncthreads();
        if (tid == 0) {
            tma_store_1d(img + size_t(t) * p.tile_bytes, tile_ptr(cur), p.tile_bytes);
            tma_commit();
        }
        n_stores += 1;
        last_store[cur] = n_stores;
    };

    const uint8_t* first_img = p.arena + ((p.flags & PF_READ_LIVE) ? size_t(0) : (size_t(p.ops[0].image_off256) << 8));
    const uint32_t first_rows = (p.flags & PF_READ_LIVE) ? p.live_rows : p.ops[0].n_rows;

    for (uint32_t tile = blockIdx.x; tile < p.n_tiles;) {
        __syncthreads();  // every thread has read the previous s_next
        if (tid == 0) s_next = gridDim.x + atomicAdd(&p.ticket[1], 1u);  // the block's next tile (read after the barrier in load_tile)
        load_tile(first_img, tile, first_rows);
        const uint32_t next_tile = s_next;

        for (uint32_t i = (p.flags & PF_READ_LIVE) ? 0u : 1u; i < p.n_ops; ++i) {
            const Op& op = p.ops[i];
            if (op.kind == OP_ADVANCE) {
                const uint32_t nb = cur ^ 1u;
                before_write(nb);
                const float dt = __uint_as_float(op.dt_bits);
                const uint8_t* al_cur = alive_ptr(cur);
                uint8_t* al_new = alive_ptr(nb);
#pragma unroll 1
                for (int k = 0; k < kGenericRowsPerThread; ++k) {
                    const uint32_t r = tid + k * kGenericBlock;
                    const uint32_t m = al_cur[r];   // the schedule's systems all see the entity as it was before the frame:
                    bool kill = false;              // despawn commands are applied after the last system
                    const uint32_t* old_row = reinterpret_cast<const uint32_t*>(tile_ptr(cur)) + r;
                    uint32_t* row = reinterpret_cast<uint32_t*>(tile_ptr(nb)) + r;
                    for (uint32_t w = 0; w < p.words; ++w) row[size_t(w) * kTileRows] = old_row[size_t(w) * kTileRows];
                    for (uint32_t s = 0; s < p.n_sys; ++s) {
                        const SysSpec sy = p.sys[s];
                        if (!row_matches(m, sy.need)) continue;
                        uint32_t* w0 = row + size_t(sy.plane0) * kTileRows;
                        switch (sy.id) {
                        case BGR_SYS_U32_ADD: w0[0] += sy.param; break;
                        case BGR_SYS_U32_SATSUB_DESPAWN: {
                            uint32_t v = w0[0];
                            v = v > sy.param ? v - sy.param : 0u;
                            w0[0] = v;
                            kill = kill || v == 0u;
                            break;
                        }
                        case BGR_SYS_U32_STORE_CALL_COUNT: w0[0] = op.call_count + sy.param; break;
                        case BGR_SYS_DESPAWN_ON_INPUT: {  // param = player handle | value << 8
                            const uint32_t player = sy.param & 0xFFu, n_players = (op.flags >> 8) & 0xFu;
                            const uint32_t input = player < n_players && player < 8 ? op.inputs[player] : 0u;
                            kill = kill || input == (sy.param >> 8);
                            break;
                        }
                        case BGR_SYS_PARTICLES_UPDATE: {
                            uint32_t* v = row + size_t(sy.plane1) * kTileRows;
                            uint32_t tx = w0[0], ty = w0[kTileRows], tz = w0[2 * kTileRows], vx = v[0], vy = v[kTileRows], vz = v[2 * kTileRows];
                            particle_step(tx, ty, tz, vx, vy, vz, dt);
                            w0[0] = tx; w0[kTileRows] = ty; w0[2 * kTileRows] = tz; v[0] = vx; v[kTileRows] = vy; v[2 * kTileRows] = vz;
                            break;
                        }
                        case BGR_SYS_PARTICLES_DESPAWN: {
                            uint64_t ttl = (uint64_t(w0[kTileRows]) << 32) | w0[0];
                            ttl -= 1;
                            w0[0] = uint32_t(ttl); w0[kTileRows] = uint32_t(ttl >> 32);
                            kill = kill || ttl == 0;
                            break;
                        }
                        case BGR_SYS_BOX_MOVE: {
                            float* t = reinterpret_cast<float*>(w0);
                            float* v = reinterpret_cast<float*>(row + size_t(sy.plane1) * kTileRows);
                            float tx = t[0], ty = t[kTileRows], tz = t[2 * kTileRows], vx = v[0], vy = v[kTileRows], vz = v[2 * kTileRows];
                            const unsigned long long handle = p.order_base + size_t(tile) * kTileRows + r;
                            const uint32_t n_players = (op.flags >> 8) & 0xFu;
                            const uint32_t input = handle < n_players && handle < 8 ? op.inputs[handle] : 0u;
                            box_move_step(tx, ty, tz, vx, vy, vz, dt, input);
                            t[0] = tx; t[kTileRows] = ty; t[2 * kTileRows] = tz; v[0] = vx; v[kTileRows] = vy; v[2 * kTileRows] = vz;
                            break;
                        }
                        default: break;
                        }
                    }
                    al_new[r] = kill ? uint8_t(0) : uint8_t(m);
                }
                cur = nb;
            } else if (op.kind == OP_SAVE) {
                // the bulk store streams the tile to the frame's slot while the threads hash their rows out of it
                if (!(op.flags & OPF_NO_STORE)) store_tile(p.arena + (size_t(op.image_off256) << 8), tile);
                const uint8_t* al = alive_ptr(cur);
                uint32_t m[kGenericRowsPerThread];
                uint64_t t0[kGenericRowsPerThread];
                uint32_t n_alive = 0, bad = 0;
#pragma unroll
                for (int k = 0; k < kGenericRowsPerThread; ++k) {
                    const uint32_t r = tid + k * kGenericBlock;
                    m[k] = al[r];
                    n_alive += m[k] & 1u;
                    t0[k] = sea_order_lane(p.order_base + size_t(tile) * kTileRows + r);
                }
                const unsigned full = 0xffffffffu;
                unsigned int* a = &s_acc[op.save_index * kAccStride * 2];
#pragma unroll 1
                for (uint32_t c = 0; c < p.n_hash; ++c) {  // one checksummed column at a time: one copy of the hash code
                    const HashSpec hs = p.hash[c];
                    uint64_t hx = 0;
#pragma unroll
                    for (int k = 0; k < kGenericRowsPerThread; ++k) {
                        if (!row_matches(m[k], hs.absent)) continue;  // Query<(&RollbackId, &T)>: exists and has the component
                        const uint32_t* col = reinterpret_cast<const uint32_t*>(tile_ptr(cur)) + size_t(hs.first_plane) * kTileRows + tid + k * kGenericBlock;
                        if (hs.finite)
                            for (uint32_t q = 0; q + 4 <= hs.len; q += 4) bad |= f32_bits_nonfinite(col[size_t((hs.off + q) >> 2) * kTileRows]);
                        hx ^= sea_hash_entity(t0[k], hash_row_range(col, hs.off, hs.len));
                    }
                    const uint32_t lo = __reduce_xor_sync(full, uint32_t(hx)), hi = __reduce_xor_sync(full, uint32_t(hx >> 32));
                    if (lane == 0) { atomicXor(&a[2 * hs.slot], lo); atomicXor(&a[2 * hs.slot + 1], hi); }
                }
                const uint32_t cnt = __reduce_add_sync(full, n_alive);
                const uint32_t anybad = __reduce_or_sync(full, bad);
                if (lane == 0) { atomicAdd(&a[12], cnt); if (anybad) atomicOr(&a[14], 1u); }
            } else {  // OP_LOAD
                load_tile(p.arena + (size_t(op.image_off256) << 8), tile, op.n_rows);
            }
        }
        if (p.flags & PF_WRITE_LIVE_ACTIVE) store_tile(p.arena, tile);
        tile = next_tile;
    }
    if (tid == 0) tma_wait_all();  // every bulk store has landed before the results are published

    // ---- block partials -> global accumulators -> (last block) host-visible results: k_particles_program's protocol ----
    __syncthreads();
    for (uint32_t i = tid; i < p.n_saves * kAccStride; i += kGenericBlock) {
        unsigned long long v = (unsigned long long)s_acc[2 * i] | ((unsigned long long)s_acc[2 * i + 1] << 32);
        const uint32_t c = i % kAccStride;
        if (v) {
            if (c == 6) atomicAdd(&p.accum[i], v);
            else if (c == 7) atomicOr(&p.accum[i], v);
            else atomicXor(&p.accum[i], v);
        }
    }
    __threadfence();
    __syncthreads();
    if (p.trace && tid == 0) atomicMax(&p.trace[1], globaltimer_ns());
    if (tid == 0) s_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (uint32_t i = tid; i < p.n_saves * kAccStride; i += kGenericBlock)
            publish_pair(p.out, i, atomicExch(&p.accum[i], 0ULL), p.seq);
        if (tid == 0) publish_pair(p.out, kSeqIndex, p.seq, p.seq);
        __syncthreads();
        if (tid == 0) {
            p.ticket[0] = 0u;
            p.ticket[1] = 0u;
            if (p.trace) p.trace[2] = globaltimer_ns();
        }
    }
}

}  // namespace bgr
