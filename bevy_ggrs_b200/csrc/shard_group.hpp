// Shard group: the cross-shard step of the hot path, inside the engine.
//
// Entity-range shards (one process + one engine per GPU, one node) never exchange state.  Per SaveGameState every
// shard produces 64 bytes of raw partials — per checksummed column the XOR over its live rows of the per-entity
// hash (component_checksum.rs:81-90, before the final `result.hash()` at :93) and its live-row count — and the
// frame checksum GGRS needs (`cell.save(frame, None, checksum)`, schedule_systems.rs:231-236) is
//     seahash(active_sum, total_sum) ^ XOR_c seahash(XOR_shards xor_c)            (entity_checksum.rs:35-43, checksum.rs:88-99)
// The consumer of that value is the HOST of every rank, so the exchange medium is host memory that all ranks map:
// one POSIX shared-memory segment, page-locked and mapped into every rank's GPU address space (cudaHostRegister).
// The last block of a rank's fused kernel already publishes its result words with plain stores to host-mapped memory,
// each as a self-validating pair (v, v ^ tag(seq, i)) (kernels.cuh publish_pair); in a group those stores land in the
// shared segment, every rank's CPU polls the pairs of all ranks' blocks and folds.  No extra kernel, no copy, no collective library call, no
// Python between two ticks: the exchange costs what the single-GPU completion poll costs.  (NCCL has no XOR
// reduction, and an all_gather of 64 B per frame through NCCL + a D2H copy is ~40 us of launch latency per tick —
// the round-1 design, which capped weak scaling at 0.78.)
//
// This header is pure host code (no CUDA): the engine registers the block area with the GPU, and the CPU-only
// tests drive the same join / publish / collect logic through the bgr_group_* entry points with a stand-in for the
// kernel's publish.
#pragma once
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/bevy_ggrs_b200.h"

namespace bgr {

constexpr uint32_t kGroupMagic = 0x47524742u;  // "BGRG"
constexpr uint32_t kGroupBufs = 8;             // result buffers per rank == max un-collected request vectors
constexpr uint32_t kGroupMaxRanks = 64;
constexpr uint32_t kGroupMaxSaves = 40;        // == kMaxSaves
constexpr uint32_t kGroupAccStride = 8;        // == kAccStride: [0..5] column xors, [6] active rows, [7] flags
// Result words are published as self-validating pairs (v, v ^ tag(seq, i)), kernels.cuh result_tag / publish_pair;
// pair kGroupMaxSaves * kGroupAccStride is the completion pair.
inline uint64_t group_result_tag(uint64_t seq, uint32_t i) {
    return ((seq * 0x9E3779B97F4A7C15ULL) ^ (uint64_t(i + 1) * 0xD6E8FEB86659FD93ULL)) | 1ULL;
}
inline bool group_pair_valid(const volatile uint64_t* blk, uint32_t i, uint64_t seq, uint64_t* v_out) {
    const uint64_t a = blk[2 * i], b = blk[2 * i + 1];
    if ((a ^ b) != group_result_tag(seq, i)) return false;
    *v_out = a;
    return true;
}

struct alignas(64) GroupHeader {
    std::atomic<uint32_t> magic, world, block_words, joined, left;
};
struct alignas(64) GroupRank {
    std::atomic<uint64_t> consumed;  // last group sequence number this rank has folded (its peers may reuse that buffer)
    std::atomic<uint64_t> seq_base;  // engine sequence number at join: the kernel of group sequence g tags its pairs with seq_base + g
};
struct alignas(64) GroupMeta {  // host-written at submit time, one per (rank, buffer)
    std::atomic<uint64_t> gseq;  // written last (release)
    uint32_t n_saves, n_columns;
    int32_t frames[kGroupMaxSaves];
    uint32_t totals[kGroupMaxSaves];  // RollbackOrdered::len() of the shard at each save
};

inline uint64_t now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return uint64_t(ts.tv_sec) * 1000u + uint64_t(ts.tv_nsec) / 1000000u;
}

// spin -> yield -> give up after timeout_ms
template <class Pred>
inline bool spin_until(Pred ok, uint64_t timeout_ms) {
    for (int i = 0; i < 4096; ++i) {
        if (ok()) return true;
        __builtin_ia32_pause();
    }
    const uint64_t t0 = now_ms();
    for (uint64_t n = 0;; ++n) {
        for (int i = 0; i < 256; ++i) {
            if (ok()) return true;
            __builtin_ia32_pause();
        }
        if ((n & 63u) == 63u) {
            if (now_ms() - t0 > timeout_ms) return false;
            sched_yield();
        }
    }
}

class ShardGroup {
public:
    uint32_t rank = 0, world = 0, block_words = 0;
    uint64_t timeout_ms = 60000;

    ~ShardGroup() { leave(); }

    // Every rank calls join with the same (name, world, block_words).  `name` must be unique per group instance
    // (e.g. launcher pid + port): rank 0 creates the segment, the others wait for it; once every rank has mapped it
    // rank 0 unlinks the name, so nothing outlives the processes even if one of them dies.
    bool join(const std::string& name, uint32_t rank_, uint32_t world_, uint32_t block_words_, uint64_t seq_base,
              std::string* err) {
        if (world_ == 0 || world_ > kGroupMaxRanks || rank_ >= world_) { *err = "bad rank / world size"; return false; }
        rank = rank_; world = world_; block_words = block_words_;
        shm_name_ = "/" + name;
        const size_t page = 4096;
        const size_t head = sizeof(GroupHeader) + sizeof(GroupRank) * world + sizeof(GroupMeta) * world * kGroupBufs;
        blocks_off_ = (head + page - 1) / page * page;
        const size_t blocks = size_t(world) * kGroupBufs * block_words * sizeof(uint64_t);
        size_ = blocks_off_ + (blocks + page - 1) / page * page;
        int fd = -1;
        if (rank == 0) {
            shm_unlink(shm_name_.c_str());  // a stale segment of a crashed earlier group with the same name
            fd = shm_open(shm_name_.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
            if (fd < 0) { *err = "shm_open(create) failed for " + shm_name_; return false; }
            if (ftruncate(fd, off_t(size_)) != 0) { close(fd); *err = "ftruncate failed"; return false; }
        } else {
            const uint64_t t0 = now_ms();
            for (;;) {
                fd = shm_open(shm_name_.c_str(), O_RDWR, 0600);
                if (fd >= 0) {
                    struct stat st;
                    if (fstat(fd, &st) == 0 && size_t(st.st_size) == size_) break;  // created AND sized
                    close(fd); fd = -1;
                }
                if (now_ms() - t0 > timeout_ms) { *err = "shard group: rank 0 never created " + shm_name_; return false; }
                usleep(200);
            }
        }
        void* m = mmap(nullptr, size_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (m == MAP_FAILED) { *err = "mmap failed"; return false; }
        base_ = static_cast<uint8_t*>(m);
        hdr_ = reinterpret_cast<GroupHeader*>(base_);
        ranks_ = reinterpret_cast<GroupRank*>(base_ + sizeof(GroupHeader));
        meta_ = reinterpret_cast<GroupMeta*>(base_ + sizeof(GroupHeader) + sizeof(GroupRank) * world);
        if (rank == 0) {  // fresh segment is zero-filled; publish the geometry, then the magic
            hdr_->world.store(world); hdr_->block_words.store(block_words);
            hdr_->magic.store(kGroupMagic, std::memory_order_release);
        } else {
            if (!spin_until([&] { return hdr_->magic.load(std::memory_order_acquire) == kGroupMagic; }, timeout_ms)) {
                *err = "shard group: segment never initialised"; return false;
            }
            if (hdr_->world.load() != world || hdr_->block_words.load() != block_words) {
                *err = "shard group: ranks disagree on world size / block size"; return false;
            }
        }
        ranks_[rank].consumed.store(0);
        ranks_[rank].seq_base.store(seq_base, std::memory_order_release);
        hdr_->joined.fetch_add(1, std::memory_order_acq_rel);
        if (!spin_until([&] { return hdr_->joined.load(std::memory_order_acquire) >= world; }, timeout_ms)) {
            *err = "shard group: not every rank joined"; return false;
        }
        if (rank == 0) shm_unlink(shm_name_.c_str());  // everyone holds a mapping now
        return true;
    }

    void leave() {
        if (!base_) return;
        hdr_->left.fetch_add(1);
        munmap(base_, size_);
        base_ = nullptr;
    }

    bool joined() const { return base_ != nullptr; }
    uint8_t* blocks_base() const { return base_ + blocks_off_; }                       // page-aligned: what the GPU maps
    size_t blocks_bytes() const { return size_ - blocks_off_; }
    uint64_t* block(uint32_t r, uint32_t buf) const {
        return reinterpret_cast<uint64_t*>(base_ + blocks_off_) + (size_t(r) * kGroupBufs + buf) * block_words;
    }
    static uint32_t buf_of(uint64_t gseq) { return uint32_t((gseq - 1) % kGroupBufs); }

    // before submitting group sequence g (which reuses the buffer of g - kGroupBufs): every peer has folded that one
    bool wait_reusable(uint64_t gseq, std::string* err) {
        if (gseq <= kGroupBufs) return true;
        const uint64_t need = gseq - kGroupBufs;
        for (uint32_t r = 0; r < world; ++r)
            if (!spin_until([&] { return ranks_[r].consumed.load(std::memory_order_acquire) >= need; }, timeout_ms)) {
                *err = "shard group: rank " + std::to_string(r) + " stopped collecting (timeout)"; return false;
            }
        return true;
    }

    void publish_meta(uint64_t gseq, uint32_t n_saves, uint32_t n_columns, const int32_t* frames, const uint32_t* totals) {
        GroupMeta& m = meta_[size_t(rank) * kGroupBufs + buf_of(gseq)];
        m.n_saves = n_saves; m.n_columns = n_columns;
        std::memcpy(m.frames, frames, sizeof(int32_t) * n_saves);
        std::memcpy(m.totals, totals, sizeof(uint32_t) * n_saves);
        m.gseq.store(gseq, std::memory_order_release);
    }

    // Wait for every rank's result block of group sequence g, check that all shards executed the same request
    // vector, and combine: XOR the column words, sum active / total, OR the flags.  out[k] gets the combined partial
    // of save k; *flags_out the OR of the non-finite flags.  `own_done` tells that the caller has already waited for
    // its own block (the engine polls / synchronises its stream itself).
    bool combine(uint64_t gseq, bgr_partial* out, uint32_t cap, uint32_t* n_out, uint64_t* flags_out, std::string* err) {
        const uint32_t buf = buf_of(gseq);
        const GroupMeta& mine = meta_[size_t(rank) * kGroupBufs + buf];
        const uint32_t n_saves = mine.n_saves;
        if (n_out) *n_out = n_saves;
        uint64_t flags = 0;
        for (uint32_t k = 0; k < n_saves && k < cap; ++k) {
            std::memset(&out[k], 0, sizeof(bgr_partial));
            out[k].frame = mine.frames[k];
            out[k].n_columns = mine.n_columns;
        }
        for (uint32_t r = 0; r < world; ++r) {
            const GroupMeta& m = meta_[size_t(r) * kGroupBufs + buf];
            const volatile uint64_t* blk = block(r, buf);
            const uint64_t want = ranks_[r].seq_base.load(std::memory_order_acquire) + gseq;
            const uint32_t seq_index = kGroupMaxSaves * kGroupAccStride;
            uint64_t words[kGroupMaxSaves * kGroupAccStride];
            uint32_t valid = 0;  // pairs [0, valid) of this rank's block have been accepted
            bool seq_ok = false;
            auto ready = [&] {
                if (m.gseq.load(std::memory_order_acquire) != gseq) return false;
                uint64_t v;
                if (!seq_ok) { if (!group_pair_valid(blk, seq_index, want, &v)) return false; seq_ok = true; }
                const uint32_t n = m.n_saves * kGroupAccStride;
                while (valid < n) {
                    if (!group_pair_valid(blk, valid, want, &words[valid])) return false;
                    ++valid;
                }
                return true;
            };
            if (!spin_until(ready, timeout_ms)) {
                *err = "shard group: rank " + std::to_string(r) + " did not publish request vector " + std::to_string(gseq) + " (timeout)";
                return false;
            }
            std::atomic_thread_fence(std::memory_order_acquire);
            if (m.n_saves != n_saves || m.n_columns != mine.n_columns ||
                std::memcmp(m.frames, mine.frames, sizeof(int32_t) * n_saves) != 0) {
                *err = "shard group: shards executed different request vectors (every shard must be handed the same Vec<GgrsRequest>)";
                return false;
            }
            for (uint32_t k = 0; k < n_saves && k < cap; ++k) {
                const uint64_t* row = words + size_t(k) * kGroupAccStride;
                for (uint32_t c = 0; c < BGR_MAX_CHECKSUM_COLUMNS; ++c) out[k].xor_[c] ^= row[c];
                out[k].active += row[6];
                out[k].total += m.totals[k];
                flags |= row[7];
            }
        }
        if (flags_out) *flags_out = flags;
        ranks_[rank].consumed.store(gseq, std::memory_order_release);
        return true;
    }

    // CPU stand-in for the kernel's publish (tests): write the partials of group sequence g into this rank's block
    void publish_block_from_host(uint64_t gseq, const bgr_partial* parts, uint32_t n) {
        volatile uint64_t* blk = block(rank, buf_of(gseq));
        const uint64_t seq = ranks_[rank].seq_base.load() + gseq;
        auto put = [&](uint32_t i, uint64_t v) { blk[2 * i] = v; blk[2 * i + 1] = v ^ group_result_tag(seq, i); };
        for (uint32_t k = 0; k < n; ++k) {
            for (uint32_t c = 0; c < BGR_MAX_CHECKSUM_COLUMNS; ++c) put(k * kGroupAccStride + c, parts[k].xor_[c]);
            put(k * kGroupAccStride + 6, parts[k].active);
            put(k * kGroupAccStride + 7, 0);
        }
        put(kGroupMaxSaves * kGroupAccStride, seq);
        std::atomic_thread_fence(std::memory_order_release);
    }

private:
    std::string shm_name_;
    uint8_t* base_ = nullptr;
    size_t size_ = 0, blocks_off_ = 0;
    GroupHeader* hdr_ = nullptr;
    GroupRank* ranks_ = nullptr;
    GroupMeta* meta_ = nullptr;
};

}  // namespace bgr
