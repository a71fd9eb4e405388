// Quorum data model + the two pure decision procedures of the control plane.
//
//   quorum_compute          - is there a valid quorum right now, and who is in it
//                             (reference semantics: src/lighthouse.rs:141-269)
//   compute_quorum_results  - per-rank view of a quorum: replica rank, recovery
//                             source/destinations, primary store
//                             (reference semantics: src/manager.rs:489-625)
//
// Both are side-effect free and unit-tested from Python through the bindings
// (tests/test_quorum_logic.py re-creates the reference's Rust test tables).
// Message fields mirror proto/torchft.proto:37-100 so a gRPC facade could be
// added later without touching the logic.
#pragma once

#include <cstdint>
#include <map>
#include <optional>
#include <string>
#include <vector>

#include "wire.h"

namespace tft {

struct QuorumMember {
  std::string replica_id;
  std::string address;        // ManagerServer address ("http://host:port")
  std::string store_address;  // replica group's TCPStore ("host:port")
  int64_t step = 0;
  uint64_t world_size = 0;
  bool shrink_only = false;
  std::string data;  // opaque JSON supplied through LighthouseClient.quorum(data=...)
  int64_t commit_failures = 0;

  void encode(Writer& w) const;
  static QuorumMember decode(Reader& r);
};

struct Quorum {
  int64_t quorum_id = 0;
  std::vector<QuorumMember> participants;
  int64_t created_ms = 0;  // unix epoch millis

  void encode(Writer& w) const;
  static Quorum decode(Reader& r);
};

struct LighthouseOpt {
  std::string bind = "[::]:29510";
  uint64_t min_replicas = 1;
  uint64_t join_timeout_ms = 60000;
  uint64_t quorum_tick_ms = 100;
  uint64_t heartbeat_timeout_ms = 5000;
  // First quorum gets id quorum_id_base + 1. A RESTARTED Lighthouse must not hand out ids that were
  // already used (process groups rendezvous under ".../torchft/<quorum_id>/<rank>" in long-lived
  // stores): -1 = quarter seconds since 2025-01-01 (fits the 32-bit epoch of the in-kernel flags until
  // 2059), larger than anything a previous incarnation reached unless it averaged more than four quorum
  // changes per second. 0 = the reference's behaviour.
  int64_t quorum_id_base = 0;
};

struct ParticipantDetails {
  int64_t joined_ms = 0;  // monotonic millis
  QuorumMember member;
};

// Everything quorum_compute looks at. Times are monotonic milliseconds.
struct LighthouseState {
  std::map<std::string, ParticipantDetails> participants;
  std::optional<Quorum> prev_quorum;
  int64_t quorum_id = 0;
  std::map<std::string, int64_t> heartbeats;
};

struct QuorumDecision {
  std::optional<std::vector<QuorumMember>> participants;  // set iff quorum is valid now
  std::string reason;
};

QuorumDecision quorum_compute(int64_t now_ms, const LighthouseState& state, const LighthouseOpt& opt);

// True when the ordered replica_id lists differ (reference: src/lighthouse.rs:133-138).
bool quorum_changed(const std::vector<QuorumMember>& a, const std::vector<QuorumMember>& b);

struct QuorumResult {
  int64_t quorum_id = 0;
  int64_t replica_rank = 0;
  int64_t replica_world_size = 1;
  std::string recover_src_manager_address;
  std::optional<int64_t> recover_src_replica_rank;
  std::vector<int64_t> recover_dst_replica_ranks;
  std::string store_address;
  int64_t max_step = 0;
  std::optional<int64_t> max_replica_rank;
  int64_t max_world_size = 1;
  bool heal = false;
  int64_t commit_failures = 0;
  std::vector<std::string> replica_ids;

  void encode(Writer& w) const;
  static QuorumResult decode(Reader& r);
};

// Throws RpcError(kStatusNotFound) when `replica_id` is not in the quorum.
QuorumResult compute_quorum_results(const std::string& replica_id, int64_t group_rank, const Quorum& quorum,
                                    bool init_sync);

int64_t monotonic_ms();
int64_t unix_ms();

}  // namespace tft
