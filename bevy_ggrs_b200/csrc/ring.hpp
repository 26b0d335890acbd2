// Host-side bookkeeping of the snapshot ring: which frame lives in which HBM slot.
//
// Behaviour contract = GgrsSnapshots<For, As> (reference src/snapshot/mod.rs:94-271):
// a newest-first queue of (frame -> snapshot) with
//   push(frame)     :144-178  drop every stored frame that is not older than `frame`
//                              (i32 wrap-around aware), prepend, then evict beyond `depth`
//   confirm(frame)  :182-199  drop from the old end while stored < frame
//   rollback(frame) :207-223  drop from the new end until the newest == frame, else fail with
//                              "Could not rollback to {frame}: no snapshot at that moment could be found."
//   get()/peek()    :226-240
// Only the *payload* differs: instead of owning a HashMap per frame, an entry names a slot of
// pre-allocated HBM; freed slots go back to a free list, so no allocation ever happens on the
// hot path.  All per-type rings of the reference move in lock-step (every SaveWorld pushes
// the same frame into each of them, every LoadWorld rolls each back to the same frame), so one
// ring serves every registered column.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

namespace bgr {

// Fixed capacity (kMaxSlots), trivially copyable: handle_requests validates a request vector against a COPY of
// the host state, so copying the ring must not allocate on the hot path.
class SlotRing {
public:
    static constexpr uint32_t kNoSlot = 0xFFFFFFFFu;
    static constexpr uint32_t kMaxSlots = 64;

    explicit SlotRing(uint32_t n_slots = 0) { reset(n_slots); }

    void reset(uint32_t n_slots) {
        n_slots_ = n_slots > kMaxSlots ? kMaxSlots : n_slots;
        entries_.n = 0;
        free_.n = 0;
        for (uint32_t s = n_slots_; s-- > 0;) free_.push_back(s);
        depth_ = 60;  // DEFAULT_FPS until sync_depth runs (mod.rs:112)
    }

    uint32_t depth() const { return depth_; }
    void set_depth(uint32_t d) { depth_ = d; }  // mod.rs:120-135 (no eviction until the next push)
    uint32_t len() const { return entries_.size(); }
    uint32_t n_slots() const { return n_slots_; }

    // Returns the slot that now holds `frame`, or kNoSlot if more than n_slots snapshots would
    // have to be alive at once (configuration error, reported by the caller).
    uint32_t push(int32_t frame) {
        // entries_ is oldest-first; the reference's "front" is our back()
        while (!entries_.empty() && !is_older(entries_.back().frame, frame)) release_back();
        // the entries that `while len > depth` would evict after the insert can go first:
        // the final queue is identical and their slots become reusable for this push
        while (!entries_.empty() && entries_.size() + 1 > depth_) release_front();
        if (depth_ == 0) return kNoSlot - 1;  // pushed and immediately evicted: nothing is stored
        if (free_.empty()) return kNoSlot;
        uint32_t s = free_.back();
        free_.pop_back();
        entries_.push_back({frame, s});
        return s;
    }

    void confirm(int32_t confirmed_frame) {
        while (!entries_.empty() && entries_.front().frame < confirmed_frame) release_front();
    }

    // false => the reference would panic; `error` gets the same text
    bool rollback(int32_t frame, std::string* error) {
        for (;;) {
            if (entries_.empty()) {
                if (error)
                    *error = "Could not rollback to " + std::to_string(frame) +
                             ": no snapshot at that moment could be found.";
                return false;
            }
            if (entries_.back().frame != frame) release_back();
            else return true;
        }
    }

    bool get(uint32_t* slot, std::string* error) const {
        if (entries_.empty()) {
            if (error) *error = "no snapshot available — call rollback(frame) before get()";
            return false;
        }
        *slot = entries_.back().slot;
        return true;
    }

    bool peek(int32_t frame, uint32_t* slot) const {
        for (uint32_t i = entries_.size(); i-- > 0;)  // newest first, like the reference's iter()
            if (entries_[i].frame == frame) { *slot = entries_[i].slot; return true; }
        return false;
    }

    // newest first
    void frames(std::vector<int32_t>* out) const {
        out->clear();
        for (uint32_t i = entries_.size(); i-- > 0;) out->push_back(entries_[i].frame);
    }

private:
    struct Entry { int32_t frame; uint32_t slot; };

    // "stored is strictly older than incoming" == NOT (current_after_frame || current_after_frame_wrapped), mod.rs:156-161
    static bool is_older(int32_t stored, int32_t incoming) {
        int64_t diff = int64_t(stored) - int64_t(incoming);
        uint32_t ad = uint32_t(diff < 0 ? -diff : diff);
        bool wrapped = ad > (UINT32_MAX / 2);
        bool after = stored >= incoming && !wrapped;
        bool after_wrapped = incoming >= stored && wrapped;
        return !(after || after_wrapped);
    }
    // tiny inline vector: no heap, trivially copyable
    template <class T>
    struct Small {
        std::array<T, kMaxSlots> a;
        uint32_t n = 0;
        bool empty() const { return n == 0; }
        uint32_t size() const { return n; }
        T& back() { return a[n - 1]; }
        const T& back() const { return a[n - 1]; }
        T& front() { return a[0]; }
        const T& front() const { return a[0]; }
        T& operator[](uint32_t i) { return a[i]; }
        const T& operator[](uint32_t i) const { return a[i]; }
        void push_back(const T& v) { a[n++] = v; }
        void pop_back() { --n; }
        void pop_front() { for (uint32_t i = 1; i < n; ++i) a[i - 1] = a[i]; --n; }
    };
    void release_back() { free_.push_back(entries_.back().slot); entries_.pop_back(); }
    void release_front() { free_.push_back(entries_.front().slot); entries_.pop_front(); }

    Small<Entry> entries_;  // oldest first; depth is small (<= 64), O(depth) per operation
    Small<uint32_t> free_;
    uint32_t n_slots_ = 0;
    uint32_t depth_ = 60;
};

}  // namespace bgr
