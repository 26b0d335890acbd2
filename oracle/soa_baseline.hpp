// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/seahash.hpp header).
//
// "Optimised CPU SoA bar" of BASELINE.md §2(2): the honesty check beside the faithful restatement.
// Same results as oracle::World (bit-identical checksums, tested), but with the data structures a CPU
// programmer would pick for this hot path instead of the reference's: flat columns, a ring of pre-allocated
// slots filled by memcpy, RollbackOrdered index == row, and the whole request vector executed per entity
// range on every host core (the same decomposition the GPU kernel uses).  Particles schema only.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/bevy_ggrs_b200.h"
#include "ggrs_snapshots.hpp"
#include "seahash.hpp"
#include "world.hpp"  // duration_as_secs_f32

namespace oracle {

struct SoaImage {
    std::vector<float> tf;       // n x 10
    std::vector<float> vel;      // n x 3
    std::vector<uint64_t> ttl;   // n
    std::vector<uint8_t> alive;  // n
    void resize(size_t n) { tf.resize(n * 10); vel.resize(n * 3); ttl.resize(n); alive.resize(n); }
};

struct SoaWorld {
    size_t n = 0;
    uint32_t fps = 60;
    unsigned threads = 1;
    SoaImage live;
    std::vector<SoaImage> slots;
    std::vector<uint32_t> free_slots;
    GgrsSnapshots<uint32_t> ring;  // payload = slot id
    std::vector<uint64_t> slot_elapsed;
    int32_t frame_count = 0, confirmed = 0;
    bool has_maxpred = false; uint32_t maxpred = 0;
    uint64_t elapsed_ns = 0;

    SoaWorld(size_t n_, uint32_t depth, uint32_t fps_, unsigned threads_) : n(n_), fps(fps_), threads(threads_ ? threads_ : 1) {
        live.resize(n);
        slots.resize(depth + 1);
        for (auto& s : slots) s.resize(n);
        for (uint32_t i = depth + 1; i-- > 0;) free_slots.push_back(i);
        slot_elapsed.assign(depth + 1, 0);
    }

    struct Op { uint32_t kind; uint32_t slot; float dt; uint32_t save_index; };

    uint32_t push_slot(int32_t frame) {
        // GgrsSnapshots::push with slot recycling: entries the push drops give their slot back
        std::vector<uint32_t> before(ring.snapshots.begin(), ring.snapshots.end());
        uint32_t s = free_slots.back();
        free_slots.pop_back();
        ring.push(frame, s);
        for (uint32_t b : before)
            if (std::find(ring.snapshots.begin(), ring.snapshots.end(), b) == ring.snapshots.end()) free_slots.push_back(b);
        return s;
    }
    void confirm(int32_t f) {
        std::vector<uint32_t> before(ring.snapshots.begin(), ring.snapshots.end());
        ring.confirm(f);
        for (uint32_t b : before)
            if (std::find(ring.snapshots.begin(), ring.snapshots.end(), b) == ring.snapshots.end()) free_slots.push_back(b);
    }
    void rollback(int32_t f) {
        std::vector<uint32_t> before(ring.snapshots.begin(), ring.snapshots.end());
        ring.rollback(f);
        for (uint32_t b : before)
            if (std::find(ring.snapshots.begin(), ring.snapshots.end(), b) == ring.snapshots.end()) free_slots.push_back(b);
    }

    void run_range(const std::vector<Op>& ops, size_t a, size_t b, uint64_t* partial /* n_saves x 3 */) {
        const size_t m = b - a;
        for (const Op& op : ops) {
            if (op.kind == BGR_REQ_LOAD) {
                const SoaImage& s = slots[op.slot];
                std::memcpy(&live.tf[a * 10], &s.tf[a * 10], m * 40);
                std::memcpy(&live.vel[a * 3], &s.vel[a * 3], m * 12);
                std::memcpy(&live.ttl[a], &s.ttl[a], m * 8);
                std::memcpy(&live.alive[a], &s.alive[a], m);
            } else if (op.kind == BGR_REQ_ADVANCE) {
                const float dt = op.dt, gx = 0.0f * 200.0f, gy = -1.0f * 200.0f, gz = 0.0f * 200.0f;
                for (size_t r = a; r < b; ++r) {
                    if (!live.alive[r]) continue;
                    float* v = &live.vel[r * 3]; float* t = &live.tf[r * 10];
                    float ax = gx * dt, ay = gy * dt, az = gz * dt;
                    v[0] = v[0] + ax; v[1] = v[1] + ay; v[2] = v[2] + az;
                    float dx = v[0] * dt, dy = v[1] * dt, dz = v[2] * dt;
                    t[0] = t[0] + dx; t[1] = t[1] + dy; t[2] = t[2] + dz;
                    live.ttl[r] -= 1;
                    if (live.ttl[r] == 0) live.alive[r] = 0;
                }
            } else {  // SAVE: checksum partials + memcpy into the slot
                uint64_t xt = 0, xv = 0, cnt = 0;
                for (size_t r = a; r < b; ++r) {
                    if (!live.alive[r]) continue;
                    ++cnt;
                    const float* t = &live.tf[r * 10]; const float* v = &live.vel[r * 3];
                    if (!std::isfinite(t[0]) || !std::isfinite(t[1]) || !std::isfinite(t[2]) || !std::isfinite(v[0]) || !std::isfinite(v[1]) || !std::isfinite(v[2]))
                        throw NonFinitePanic("Hashing is not stable for NaN f32 values.");
                    SeaHasher ht; ht.write_u64(uint64_t(r)); ht.write_u64(seahash(t, 12)); xt ^= ht.finish();
                    SeaHasher hv; hv.write_u64(uint64_t(r)); hv.write_u64(seahash(v, 12)); xv ^= hv.finish();
                }
                partial[op.save_index * 3 + 0] ^= xv;   // Velocity is registered first (checksum column 0)
                partial[op.save_index * 3 + 1] ^= xt;
                partial[op.save_index * 3 + 2] += cnt;
                SoaImage& s = slots[op.slot];
                std::memcpy(&s.tf[a * 10], &live.tf[a * 10], m * 40);
                std::memcpy(&s.vel[a * 3], &live.vel[a * 3], m * 12);
                std::memcpy(&s.ttl[a], &live.ttl[a], m * 8);
                std::memcpy(&s.alive[a], &live.alive[a], m);
            }
        }
    }

    void handle_requests(const bgr_session_info& sess, const bgr_request* reqs, uint32_t nreq, std::vector<bgr_checksum>& out) {
        std::vector<Op> ops;
        std::vector<int32_t> save_frames;
        for (uint32_t i = 0; i < nreq; ++i) {  // host bookkeeping exactly like handle_requests (schedule_systems.rs:189-270)
            const bgr_request& rq = reqs[i];
            int32_t current = frame_count;
            if (sess.kind == BGR_SESSION_SYNCTEST) { has_maxpred = true; maxpred = sess.max_prediction; int32_t cf = current - int32_t(sess.check_distance); if (cf >= 0) confirmed = cf; }
            else if (sess.kind == BGR_SESSION_P2P) { has_maxpred = true; maxpred = sess.max_prediction; confirmed = sess.confirmed_frame; }
            if (rq.kind == BGR_REQ_SAVE) {
                if (has_maxpred) ring.set_depth(maxpred);
                confirm(confirmed);
                uint32_t s = push_slot(frame_count);
                slot_elapsed[s] = elapsed_ns;
                ops.push_back({BGR_REQ_SAVE, s, 0.f, uint32_t(save_frames.size())});
                save_frames.push_back(rq.frame);
            } else if (rq.kind == BGR_REQ_LOAD) {
                frame_count = rq.frame;
                rollback(rq.frame);
                uint32_t s = ring.get();
                elapsed_ns = slot_elapsed[s];
                ops.push_back({BGR_REQ_LOAD, s, 0.f, 0});
            } else {
                frame_count += 1;
                uint64_t runtime = uint64_t(int64_t(frame_count)) * 1000000000ULL / fps;
                float dt = duration_as_secs_f32(runtime - elapsed_ns);
                elapsed_ns = runtime;
                ops.push_back({BGR_REQ_ADVANCE, 0, dt, 0});
            }
        }
        const size_t ns = save_frames.size();
        std::vector<std::vector<uint64_t>> partials(threads, std::vector<uint64_t>(ns * 3 + 1, 0));
        std::vector<std::thread> pool;
        std::vector<std::string> errors(threads);
        for (unsigned t = 0; t < threads; ++t) {
            size_t a = n * t / threads, b = n * (t + 1) / threads;
            auto work = [this, &ops, a, b, &partials, &errors, t] {
                try { run_range(ops, a, b, partials[t].data()); } catch (const std::exception& e) { errors[t] = e.what(); }
            };
            if (t + 1 < threads) pool.emplace_back(work); else work();
        }
        for (auto& th : pool) th.join();
        for (auto& e : errors) if (!e.empty()) throw NonFinitePanic(e);
        for (size_t k = 0; k < ns; ++k) {
            uint64_t xv = 0, xt = 0, cnt = 0;
            for (unsigned t = 0; t < threads; ++t) { xv ^= partials[t][k * 3]; xt ^= partials[t][k * 3 + 1]; cnt += partials[t][k * 3 + 2]; }
            SeaHasher he; he.write_u64(cnt); he.write_u64(uint64_t(n));
            SeaHasher hv; hv.write_u64(xv);
            SeaHasher ht; ht.write_u64(xt);
            out.push_back(bgr_checksum{save_frames[k], 1u, he.finish() ^ hv.finish() ^ ht.finish(), 0});
        }
    }
};

}  // namespace oracle
