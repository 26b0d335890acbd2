// Stand-in for the ggrs sessions that feed handle_requests (C++ twin of bevy_ggrs_b200/session.py).
//
// ggrs — input queues, prediction, UDP — is third-party and OUT OF SCOPE.  The hot
// path only consumes the request vector `session.advance_frame()` returns (reference
// src/schedule_systems.rs:98,156), so only that is restated: the order of Save / Load / Advance
// requests of ggrs 0.11 `SyncTestSession::advance_frame` + `adjust_gamestate`, and the checksum
// comparison that yields `GgrsError::MismatchedChecksum` (-> SyncTestMismatch, lib.rs:131-137).
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <optional>
#include <stdexcept>
#include <utility>
#include <vector>

namespace ggrs {

using Frame = int32_t;
using PlayerHandle = size_t;
enum class InputStatus : uint8_t { Confirmed = 0, Predicted = 1, Disconnected = 2 };

struct GgrsRequest {
    enum Kind { SaveGameState = 0, LoadGameState = 1, AdvanceFrame = 2 } kind;
    Frame frame = 0;
    std::vector<std::pair<uint8_t, InputStatus>> inputs;
};

struct MismatchedChecksum {
    Frame current_frame;
    std::vector<Frame> mismatched_frames;
};

struct InvalidRequest : std::runtime_error { using std::runtime_error::runtime_error; };

class SyncTestSession {
public:
    SyncTestSession(size_t num_players, size_t check_distance, size_t max_prediction = 8, size_t input_delay = 0)
        : num_players_(num_players), check_distance_(check_distance), max_prediction_(max_prediction),
          input_delay_(input_delay), queues_(num_players), cells_(max_prediction + 1) {
        if (check_distance >= max_prediction) throw InvalidRequest("Check distance too big.");
    }
    size_t num_players() const { return num_players_; }
    size_t max_prediction() const { return max_prediction_; }
    size_t check_distance() const { return check_distance_; }
    Frame current_frame() const { return current_frame_; }

    void add_local_input(PlayerHandle handle, uint8_t input) {
        if (handle >= num_players_) throw InvalidRequest("The player handle you provided is not valid.");
        local_[handle] = input;
    }
    // GameStateCell::save(frame, None, checksum)
    void save_cell(Frame frame, std::optional<unsigned __int128> checksum) {
        cells_[size_t(frame) % cells_.size()] = std::make_pair(frame, checksum);
    }
    // Ok(requests) or Err(MismatchedChecksum)
    bool advance_frame(std::vector<GgrsRequest>& requests, MismatchedChecksum& err) {
        requests.clear();
        const Frame d = Frame(check_distance_), cur = current_frame_;
        if (d > 0 && cur > d) {
            std::vector<Frame> bad;
            for (Frame f = cur - d; f <= cur; ++f)
                if (!checksums_consistent(f)) bad.push_back(f);
            if (!bad.empty()) { err = MismatchedChecksum{cur, bad}; return false; }
            const Frame frame_to = cur - d;  // adjust_gamestate
            requests.push_back({GgrsRequest::LoadGameState, frame_to, {}});
            current_frame_ = frame_to;
            for (Frame i = 0; i < d; ++i) {
                if (i > 0) requests.push_back({GgrsRequest::SaveGameState, current_frame_, {}});
                requests.push_back(advance_request());
                ++current_frame_;
            }
        }
        if (local_.size() != num_players_) throw InvalidRequest("Missing local input while calling advance_frame().");
        for (auto& kv : local_) queues_[kv.first][current_frame_ + Frame(input_delay_)] = kv.second;
        local_.clear();
        if (d > 0) requests.push_back({GgrsRequest::SaveGameState, current_frame_, {}});
        requests.push_back(advance_request());
        ++current_frame_;
        return true;
    }

private:
    GgrsRequest advance_request() const {
        GgrsRequest r{GgrsRequest::AdvanceFrame, 0, {}};
        for (size_t p = 0; p < num_players_; ++p) {
            auto it = queues_[p].find(current_frame_);
            r.inputs.emplace_back(it == queues_[p].end() ? uint8_t(0) : it->second, InputStatus::Confirmed);
        }
        return r;
    }
    bool checksums_consistent(Frame frame_to_check) {
        const Frame oldest = current_frame_ - Frame(check_distance_);
        for (auto it = history_.begin(); it != history_.end();)
            it = it->first < oldest ? history_.erase(it) : std::next(it);
        auto& cell = cells_[size_t(frame_to_check) % cells_.size()];
        if (!cell || cell->first != frame_to_check) return true;
        auto h = history_.find(cell->first);
        if (h != history_.end()) return h->second == cell->second;
        history_[cell->first] = cell->second;
        return true;
    }

    size_t num_players_, check_distance_, max_prediction_, input_delay_;
    Frame current_frame_ = 0;
    std::vector<std::map<Frame, uint8_t>> queues_;
    std::map<PlayerHandle, uint8_t> local_;
    std::map<Frame, std::optional<unsigned __int128>> history_;
    std::vector<std::optional<std::pair<Frame, std::optional<unsigned __int128>>>> cells_;
};

// ---- trace-driven stand-ins for the two networked session kinds --------------------------------------------------
// handle_requests treats the three kinds differently only through the numbers it reads from the session per request
// (schedule_systems.rs:195-220): P2P -> (max_prediction, confirmed_frame), Spectator -> max_prediction 0 and
// confirmed = current frame.  The sockets, input exchange and prediction of ggrs stay out of scope; what is restated
// is the SHAPE of the request vectors those sessions emit, driven by a script instead of a remote peer:
//   P2P        per tick a rollback depth L (0 = the remote input arrived in time): [Load(f-L), (Adv, Save) x (L-1), Adv,] Save(f), Adv
//              confirmed_frame() = the last frame for which every "remote" input has arrived (scripted lag)
//   Spectator  per tick k >= 0 confirmed frames received from the host: Adv x k (k > 1 when catching up), never Save / Load
enum class SessionState { Synchronizing, Running };

class P2PTraceSession {
public:
    P2PTraceSession(size_t num_players, size_t max_prediction, std::vector<int> rollback_depths, int confirm_lag = 2, size_t input_delay = 0)
        : num_players_(num_players), max_prediction_(max_prediction), depths_(std::move(rollback_depths)), confirm_lag_(confirm_lag),
          input_delay_(input_delay), queues_(num_players) {}
    size_t num_players() const { return num_players_; }
    size_t max_prediction() const { return max_prediction_; }
    std::vector<PlayerHandle> local_player_handles() const { return {0}; }
    SessionState current_state() const { return SessionState::Running; }
    Frame current_frame() const { return current_frame_; }
    Frame confirmed_frame() const { return current_frame_ - Frame(confirm_lag_) < 0 ? -1 : current_frame_ - Frame(confirm_lag_); }
    int frames_ahead() const { return 0; }
    void poll_remote_clients() {}
    void add_local_input(PlayerHandle handle, uint8_t input) { local_[handle] = input; }
    void save_cell(Frame, std::optional<unsigned __int128>) {}
    std::vector<GgrsRequest> advance_frame() {
        std::vector<GgrsRequest> requests;
        int depth = tick_ < depths_.size() ? depths_[tick_] : 0;
        ++tick_;
        depth = std::min<int>({depth, int(current_frame_), int(max_prediction_), confirm_lag_});  // ggrs never rolls back past a confirmed frame
        if (depth > 0) {
            current_frame_ -= depth;
            requests.push_back({GgrsRequest::LoadGameState, current_frame_, {}});
            for (int i = 0; i < depth; ++i) {
                if (i > 0) requests.push_back({GgrsRequest::SaveGameState, current_frame_, {}});
                requests.push_back(advance_request(InputStatus::Confirmed));
                ++current_frame_;
            }
        }
        for (auto& kv : local_) queues_[kv.first][current_frame_ + Frame(input_delay_)] = kv.second;
        local_.clear();
        requests.push_back({GgrsRequest::SaveGameState, current_frame_, {}});
        requests.push_back(advance_request(InputStatus::Predicted));
        ++current_frame_;
        return requests;
    }

private:
    GgrsRequest advance_request(InputStatus remote) const {
        GgrsRequest r{GgrsRequest::AdvanceFrame, 0, {}};
        for (size_t p = 0; p < num_players_; ++p) {
            auto it = queues_[p].find(current_frame_);
            r.inputs.emplace_back(it == queues_[p].end() ? uint8_t(0) : it->second, p == 0 ? InputStatus::Confirmed : remote);
        }
        return r;
    }
    size_t num_players_, max_prediction_;
    std::vector<int> depths_;
    int confirm_lag_;
    size_t input_delay_, tick_ = 0;
    Frame current_frame_ = 0;
    std::vector<std::map<Frame, uint8_t>> queues_;
    std::map<PlayerHandle, uint8_t> local_;
};

class SpectatorTraceSession {
public:
    SpectatorTraceSession(size_t num_players, std::vector<int> frames_per_tick) : num_players_(num_players), script_(std::move(frames_per_tick)) {}
    size_t num_players() const { return num_players_; }
    SessionState current_state() const { return SessionState::Running; }
    Frame current_frame() const { return current_frame_; }
    void poll_remote_clients() {}
    // Ok(requests); an empty vector models GgrsError::PredictionThreshold ("Waiting for input from host")
    std::vector<GgrsRequest> advance_frame() {
        std::vector<GgrsRequest> requests;
        const int k = tick_ < script_.size() ? script_[tick_] : 1;
        ++tick_;
        for (int i = 0; i < k; ++i) {
            GgrsRequest r{GgrsRequest::AdvanceFrame, 0, {}};
            for (size_t p = 0; p < num_players_; ++p) r.inputs.emplace_back(uint8_t((current_frame_ + Frame(p)) & 3), InputStatus::Confirmed);
            requests.push_back(r);
            ++current_frame_;
        }
        return requests;
    }

private:
    size_t num_players_, tick_ = 0;
    std::vector<int> script_;
    Frame current_frame_ = 0;
};

}  // namespace ggrs
