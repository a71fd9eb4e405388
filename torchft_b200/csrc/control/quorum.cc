// Pure decision procedures + message codecs. Reference semantics: quorum_compute src/lighthouse.rs:141-269,
// quorum_changed :133-138, compute_quorum_results src/manager.rs:489-625 (recovery assignment: up-to-date
// replicas serve the stragglers round-robin, offset by group rank; primary store = group_rank % n-th
// up-to-date replica; init_sync forces everybody but the primary to heal at step 0).
#include "quorum.h"

#include <algorithm>
#include <chrono>
#include <set>
#include <sstream>

namespace tft {

int64_t monotonic_ms() {
  return std::chrono::duration_cast<Millis>(Clock::now().time_since_epoch()).count();
}
int64_t unix_ms() {
  return std::chrono::duration_cast<Millis>(std::chrono::system_clock::now().time_since_epoch()).count();
}

void QuorumMember::encode(Writer& w) const {
  w.str(replica_id);
  w.str(address);
  w.str(store_address);
  w.i64(step);
  w.u64(world_size);
  w.boolean(shrink_only);
  w.str(data);
  w.i64(commit_failures);
}
QuorumMember QuorumMember::decode(Reader& r) {
  QuorumMember m;
  m.replica_id = r.str();
  m.address = r.str();
  m.store_address = r.str();
  m.step = r.i64();
  m.world_size = r.u64();
  m.shrink_only = r.boolean();
  m.data = r.str();
  m.commit_failures = r.i64();
  return m;
}

void Quorum::encode(Writer& w) const {
  w.i64(quorum_id);
  w.i64(created_ms);
  w.u32((uint32_t)participants.size());
  for (const auto& p : participants) p.encode(w);
}
Quorum Quorum::decode(Reader& r) {
  Quorum q;
  q.quorum_id = r.i64();
  q.created_ms = r.i64();
  uint32_t n = r.u32();
  q.participants.reserve(n);
  for (uint32_t i = 0; i < n; ++i) q.participants.push_back(QuorumMember::decode(r));
  return q;
}

void QuorumResult::encode(Writer& w) const {
  w.i64(quorum_id);
  w.i64(replica_rank);
  w.i64(replica_world_size);
  w.str(recover_src_manager_address);
  w.boolean(recover_src_replica_rank.has_value());
  w.i64(recover_src_replica_rank.value_or(0));
  w.u32((uint32_t)recover_dst_replica_ranks.size());
  for (auto v : recover_dst_replica_ranks) w.i64(v);
  w.str(store_address);
  w.i64(max_step);
  w.boolean(max_replica_rank.has_value());
  w.i64(max_replica_rank.value_or(0));
  w.i64(max_world_size);
  w.boolean(heal);
  w.i64(commit_failures);
  w.u32((uint32_t)replica_ids.size());
  for (const auto& s : replica_ids) w.str(s);
}
QuorumResult QuorumResult::decode(Reader& r) {
  QuorumResult q;
  q.quorum_id = r.i64();
  q.replica_rank = r.i64();
  q.replica_world_size = r.i64();
  q.recover_src_manager_address = r.str();
  bool has = r.boolean();
  int64_t v = r.i64();
  if (has) q.recover_src_replica_rank = v;
  uint32_t n = r.u32();
  for (uint32_t i = 0; i < n; ++i) q.recover_dst_replica_ranks.push_back(r.i64());
  q.store_address = r.str();
  q.max_step = r.i64();
  has = r.boolean();
  v = r.i64();
  if (has) q.max_replica_rank = v;
  q.max_world_size = r.i64();
  q.heal = r.boolean();
  q.commit_failures = r.i64();
  n = r.u32();
  for (uint32_t i = 0; i < n; ++i) q.replica_ids.push_back(r.str());
  return q;
}

bool quorum_changed(const std::vector<QuorumMember>& a, const std::vector<QuorumMember>& b) {
  if (a.size() != b.size()) return true;
  for (size_t i = 0; i < a.size(); ++i)
    if (a[i].replica_id != b[i].replica_id) return true;
  return false;
}

QuorumDecision quorum_compute(int64_t now_ms, const LighthouseState& state, const LighthouseOpt& opt) {
  // 1. who is alive: heartbeat strictly younger than the timeout
  std::set<std::string> healthy_replicas;
  for (const auto& [id, last] : state.heartbeats)
    if (now_ms - last < (int64_t)opt.heartbeat_timeout_ms) healthy_replicas.insert(id);

  // participants that are also alive; std::map iteration == sorted by replica_id
  std::vector<const ParticipantDetails*> healthy_participants;
  for (const auto& [id, det] : state.participants)
    if (healthy_replicas.count(id)) healthy_participants.push_back(&det);

  std::vector<QuorumMember> candidates;
  bool shrink_only = false;
  for (const auto* d : healthy_participants) {
    candidates.push_back(d->member);
    shrink_only = shrink_only || d->member.shrink_only;
  }

  std::ostringstream md;
  md << "[" << healthy_participants.size() << "/" << state.participants.size() << " participants healthy]["
     << healthy_replicas.size() << " heartbeating][shrink_only=" << (shrink_only ? "true" : "false") << "]";
  const std::string metadata = md.str();

  // 2./3. relation to the previous quorum
  if (state.prev_quorum.has_value()) {
    std::set<std::string> prev_ids;
    for (const auto& p : state.prev_quorum->participants) prev_ids.insert(p.replica_id);
    if (shrink_only) {
      candidates.erase(std::remove_if(candidates.begin(), candidates.end(),
                                      [&](const QuorumMember& m) { return !prev_ids.count(m.replica_id); }),
                       candidates.end());
    }
    // fast quorum: every previous member is back (and healthy) -> go now, possibly grown
    bool fast = true;
    for (const auto& id : prev_ids) {
      bool found = false;
      for (const auto* d : healthy_participants)
        if (d->member.replica_id == id) {
          found = true;
          break;
        }
      if (!found) {
        fast = false;
        break;
      }
    }
    if (fast) return {std::move(candidates), "Fast quorum found! " + metadata};
  }

  // 4. minimum size
  if (healthy_participants.size() < opt.min_replicas) {
    std::ostringstream s;
    s << "New quorum not ready, only have " << healthy_participants.size() << " participants, need min_replicas "
      << opt.min_replicas << " " << metadata;
    return {std::nullopt, s.str()};
  }
  // 5. split-brain guard: strictly more than half of everything that heartbeats
  if (healthy_participants.size() <= healthy_replicas.size() / 2) {
    std::ostringstream s;
    s << "New quorum not ready, only have " << healthy_participants.size()
      << " participants, need at least half of " << healthy_replicas.size() << " healthy workers " << metadata;
    return {std::nullopt, s.str()};
  }
  // 6. stragglers: alive but not (yet) asking for this quorum
  const bool all_joined = healthy_participants.size() == healthy_replicas.size();
  int64_t first_joined = now_ms;
  for (const auto* d : healthy_participants) first_joined = std::min(first_joined, d->joined_ms);
  if (!all_joined && now_ms - first_joined < (int64_t)opt.join_timeout_ms) {
    std::ostringstream s;
    s << "Valid quorum with " << healthy_participants.size() << " participants, waiting for "
      << (healthy_replicas.size() - healthy_participants.size())
      << " healthy but not participating stragglers due to join timeout " << metadata;
    return {std::nullopt, s.str()};
  }
  return {std::move(candidates), "Valid quorum found " + metadata};
}

QuorumResult compute_quorum_results(const std::string& replica_id, int64_t group_rank, const Quorum& quorum,
                                    bool init_sync) {
  std::vector<QuorumMember> parts = quorum.participants;
  std::sort(parts.begin(), parts.end(),
            [](const QuorumMember& a, const QuorumMember& b) { return a.replica_id < b.replica_id; });
  const size_t n = parts.size();
  size_t me = n;
  for (size_t i = 0; i < n; ++i)
    if (parts[i].replica_id == replica_id) {
      me = i;
      break;
    }
  if (me == n) throw RpcError(kStatusNotFound, "replica " + replica_id + " not participating in returned quorum");

  int64_t max_step = parts[0].step;
  for (const auto& p : parts) max_step = std::max(max_step, p.step);
  std::vector<size_t> at_max;  // indices (into parts) of members at max_step
  for (size_t i = 0; i < n; ++i)
    if (parts[i].step == max_step) at_max.push_back(i);

  QuorumResult out;
  for (size_t k = 0; k < at_max.size(); ++k)
    if (at_max[k] == me) out.max_replica_rank = (int64_t)k;

  // One rendezvous store per replica group; shard rank r uses the store of the
  // (r mod |at_max|)-th up-to-date replica, spreading load across replicas.
  const size_t primary = at_max[(size_t)group_rank % at_max.size()];

  const bool force_recover = init_sync && max_step == 0;
  std::vector<size_t> recovering, up_to_date;
  for (size_t i = 0; i < n; ++i) {
    const bool behind = parts[i].step != max_step;
    const bool forced = force_recover && parts[primary].replica_id != parts[i].replica_id;
    (behind || forced ? recovering : up_to_date).push_back(i);
  }
  // round-robin assignment of sources, offset by the shard rank
  for (size_t k = 0; k < recovering.size(); ++k) {
    const size_t src = up_to_date[(k + (size_t)group_rank) % up_to_date.size()];
    if (src == me) out.recover_dst_replica_ranks.push_back((int64_t)recovering[k]);
    if (recovering[k] == me) out.recover_src_replica_rank = (int64_t)src;
  }
  out.heal = out.recover_src_replica_rank.has_value();
  if (out.heal) out.recover_src_manager_address = parts[(size_t)*out.recover_src_replica_rank].address;
  out.quorum_id = quorum.quorum_id;
  out.store_address = parts[primary].store_address;
  out.max_step = max_step;
  out.max_world_size = (int64_t)at_max.size();
  out.replica_rank = (int64_t)me;
  out.replica_world_size = (int64_t)n;
  for (const auto& p : parts) {
    out.commit_failures = std::max(out.commit_failures, p.commit_failures);
    out.replica_ids.push_back(p.replica_id);
  }
  return out;
}

}  // namespace tft
