#include "manager_server.h"

#include <unistd.h>

#include <algorithm>
#include <cstdio>

namespace tft {

static void log_replica(const std::string& replica_id, const std::string& msg) {
  static const bool quiet = [] {
    const char* e = getenv("TORCHFT_B200_LOG");
    return !(e && (std::string(e) == "info" || std::string(e) == "debug"));
  }();
  if (quiet) return;
  // "name:uuid" -> log only the human part
  const std::string shown = replica_id.substr(0, replica_id.find(':'));
  fprintf(stderr, "[torchft_b200 manager %s] %s\n", shown.c_str(), msg.c_str());
}

ManagerServer::ManagerServer(std::string replica_id, std::string lighthouse_addr, std::string hostname,
                             const std::string& bind, std::string store_addr, uint64_t world_size,
                             Millis heartbeat_interval, Millis connect_timeout, int64_t quorum_retries)
    : replica_id_(std::move(replica_id)),
      lighthouse_addr_(std::move(lighthouse_addr)),
      hostname_(std::move(hostname)),
      store_addr_(std::move(store_addr)),
      world_size_(world_size),
      heartbeat_interval_(heartbeat_interval),
      connect_timeout_(connect_timeout),
      quorum_retries_(quorum_retries) {
  // fail construction if the lighthouse is unreachable within connect_timeout
  lh_client_ = std::make_shared<LighthouseClient>(lighthouse_addr_, connect_timeout_);
  start(bind, "tft-manager");
  heartbeat_thread_ = std::thread([this] {
    name_this_thread("tft-manager-hb");
    heartbeat_loop();
  });
}

ManagerServer::~ManagerServer() { shutdown(); }

void ManagerServer::shutdown() {
  {
    std::lock_guard<std::mutex> g(mu_);
    if (shutdown_) return;
    shutdown_ = true;
  }
  cv_.notify_all();
  if (heartbeat_thread_.joinable()) heartbeat_thread_.join();
  stop();
  std::unique_lock<std::mutex> lk(mu_);
  while (quorum_workers_ != 0) {
    // unblock helpers stuck in a lighthouse long-poll
    lk.unlock();
    lighthouse_client(false)->cancel();
    lk.lock();
    cv_.wait_for(lk, Millis(100), [this] { return quorum_workers_ == 0; });
  }
}

std::string ManagerServer::address() const { return "http://" + hostname_ + ":" + std::to_string(port()); }

std::shared_ptr<LighthouseClient> ManagerServer::lighthouse_client(bool reconnect) {
  if (reconnect) {
    // the lighthouse may have restarted: build a fresh client outside the lock
    std::shared_ptr<LighthouseClient> fresh;
    try {
      fresh = std::make_shared<LighthouseClient>(lighthouse_addr_, connect_timeout_);
    } catch (const std::exception& e) {
      log_replica(replica_id_, std::string("Failed to connect to lighthouse. error: ") + e.what());
    }
    std::lock_guard<std::mutex> g(lh_mu_);
    if (fresh) lh_client_ = fresh;
    return lh_client_;
  }
  std::lock_guard<std::mutex> g(lh_mu_);
  return lh_client_;
}

void ManagerServer::heartbeat_loop() {
  std::unique_lock<std::mutex> lk(mu_);
  while (!shutdown_) {
    lk.unlock();
    try {
      lighthouse_client(false)->heartbeat(replica_id_, std::max(Millis(1000), heartbeat_interval_ * 10));
    } catch (const std::exception& e) {
      log_replica(replica_id_, std::string("Failed to send heartbeat to lighthouse: ") + e.what());
      bool stop_now;
      {
        std::lock_guard<std::mutex> g(mu_);
        stop_now = shutdown_;
      }
      if (!stop_now) lighthouse_client(true);
    }
    lk.lock();
    cv_.wait_for(lk, heartbeat_interval_, [this] { return shutdown_; });
  }
}

Quorum ManagerServer::quorum_with_retries(const QuorumMember& requester, Millis timeout) {
  int64_t retry = 0;
  while (true) {
    int64_t sleep_ms = 100;
    bool timed_out = false;
    std::string last;
    try {
      return lighthouse_client(false)->quorum(requester, timeout);
    } catch (const TimeoutError& e) {
      timed_out = true;
      last = e.what();
      log_replica(replica_id_, std::string("lighthouse quorum timeout. error: ") + e.what());
    } catch (const std::exception& e) {
      last = e.what();
      log_replica(replica_id_, std::string("lighthouse quorum failed. error: ") + e.what());
      sleep_ms = std::max<int64_t>(100, timeout.count() / std::max<int64_t>(quorum_retries_ + 1, 1));
    }
    if (retry == quorum_retries_) {
      // A deadline stays a deadline for the group's waiters (Python TimeoutError), whichever of
      // "their own wait expired" / "the forwarded request expired" is observed first.
      if (timed_out) throw TimeoutError("lighthouse quorum timed out after " + std::to_string(retry) + " retries: " + last);
      throw RpcError(kStatusInternal, "lighthouse quorum failed after " + std::to_string(retry) + " retries: " + last);
    }
    {
      std::unique_lock<std::mutex> lk(mu_);
      if (cv_.wait_for(lk, Millis(sleep_ms), [this] { return shutdown_; }))
        throw RpcError(kStatusCancelled, "manager shutting down");
    }
    lighthouse_client(true);
    ++retry;
  }
}

void ManagerServer::run_quorum(QuorumMember requester, Millis timeout, uint64_t round) {
  log_replica(replica_id_, "All workers joined - starting quorum");
  Quorum q;
  std::string err;
  uint32_t status = kStatusOk;
  try {
    q = quorum_with_retries(requester, timeout);
  } catch (const TimeoutError& e) {
    status = kStatusDeadline;
    err = e.what();
  } catch (const RpcError& e) {
    status = e.status;
    err = e.what();
  } catch (const std::exception& e) {
    status = kStatusInternal;
    err = e.what();
  }
  std::lock_guard<std::mutex> g(mu_);
  --quorum_workers_;
  // A newer round started while this one was stuck (its waiters timed out):
  // publishing now would hand the new round a stale quorum, so drop it.
  if (round == active_round_) {
    // Unlike the reference (TODO at src/manager.rs:238) failures are broadcast
    // too, so the group's waiters fail fast instead of hanging to their deadline.
    latest_quorum_ = std::move(q);
    latest_error_ = err;
    latest_error_status_ = status;
    ++quorum_gen_;
  }
  cv_.notify_all();
}

uint32_t ManagerServer::handle_rpc(uint32_t method, const std::string& req, TimePoint deadline, std::string* resp) {
  Reader r(req);
  switch (method) {
    case kManagerQuorum: {
      const int64_t group_rank = r.i64();
      const int64_t step = r.i64();
      const std::string ckpt_meta = r.str();
      const bool shrink_only = r.boolean();
      const bool init_sync = r.boolean();
      const int64_t commit_failures = r.i64();
      const Millis timeout = std::max(Millis(1), std::chrono::duration_cast<Millis>(deadline - Clock::now()));
      std::unique_lock<std::mutex> lk(mu_);
      checkpoint_metadata_[group_rank] = ckpt_meta;
      QuorumMember me;
      me.replica_id = replica_id_;
      me.address = address();
      me.store_address = store_addr_;
      me.step = step;
      me.world_size = world_size_;
      me.shrink_only = shrink_only;
      me.commit_failures = commit_failures;
      participants_[group_rank] = me;
      const uint64_t seen = quorum_gen_;
      if (participants_.size() == world_size_) {
        participants_.clear();
        const uint64_t round = ++active_round_;
        ++quorum_workers_;
        std::thread([this, me, timeout, round] {
          name_this_thread("tft-manager-q");
          run_quorum(me, timeout, round);
        }).detach();
      }
      while (quorum_gen_ <= seen) {
        if (shutdown_) {
          *resp = "manager shutting down";
          return kStatusCancelled;
        }
        if (cv_.wait_until(lk, deadline) == std::cv_status::timeout && quorum_gen_ <= seen) {
          *resp = "manager quorum timed out waiting for the replica group / lighthouse";
          return kStatusDeadline;
        }
      }
      if (latest_error_status_ != kStatusOk) {
        *resp = latest_error_;
        return latest_error_status_;
      }
      const Quorum q = latest_quorum_;
      lk.unlock();
      QuorumResult out = compute_quorum_results(replica_id_, group_rank, q, init_sync);
      Writer w;
      out.encode(w);
      *resp = w.take();
      return kStatusOk;
    }
    case kManagerCheckpointMetadata: {
      const int64_t rank = r.i64();
      std::lock_guard<std::mutex> g(mu_);
      auto it = checkpoint_metadata_.find(rank);
      if (it == checkpoint_metadata_.end()) {
        *resp = "rank not found";
        return kStatusInvalid;
      }
      Writer w;
      w.str(it->second);
      *resp = w.take();
      return kStatusOk;
    }
    case kManagerShouldCommit: {
      const int64_t group_rank = r.i64();
      (void)r.i64();  // step (not validated, as in the reference)
      const bool vote = r.boolean();
      std::unique_lock<std::mutex> lk(mu_);
      if (!vote) commit_failures_.insert(group_rank);
      commit_count_.insert(group_rank);
      const uint64_t seen = commit_gen_;
      if (commit_count_.size() == world_size_) {
        commit_decision_ = commit_failures_.empty();
        commit_count_.clear();
        commit_failures_.clear();
        ++commit_gen_;
        cv_.notify_all();
      }
      while (commit_gen_ <= seen) {
        if (shutdown_) {
          *resp = "manager shutting down";
          return kStatusCancelled;
        }
        if (cv_.wait_until(lk, deadline) == std::cv_status::timeout && commit_gen_ <= seen) {
          *resp = "should_commit timed out waiting for all group ranks";
          return kStatusDeadline;
        }
      }
      Writer w;
      w.boolean(commit_decision_);
      *resp = w.take();
      return kStatusOk;
    }
    case kManagerKill: {
      const std::string msg = r.str();
      fprintf(stderr, "[torchft_b200 manager] got kill request: %s\n", msg.c_str());
      fflush(stderr);
      _exit(1);
    }
    default:
      *resp = "unknown manager method";
      return kStatusInvalid;
  }
}

// ------------------------------------------------------------------ client
QuorumResult ManagerClient::quorum(int64_t group_rank, int64_t step, const std::string& checkpoint_metadata,
                                   bool shrink_only, Millis timeout, int64_t commit_failures, bool init_sync) {
  Writer w;
  w.i64(group_rank);
  w.i64(step);
  w.str(checkpoint_metadata);
  w.boolean(shrink_only);
  w.boolean(init_sync);
  w.i64(commit_failures);
  std::string resp = rpc_.call(kManagerQuorum, w.data(), timeout);
  Reader r(resp);
  return QuorumResult::decode(r);
}

std::string ManagerClient::checkpoint_metadata(int64_t rank, Millis timeout) {
  Writer w;
  w.i64(rank);
  std::string resp = rpc_.call(kManagerCheckpointMetadata, w.data(), timeout);
  Reader r(resp);
  return r.str();
}

bool ManagerClient::should_commit(int64_t group_rank, int64_t step, bool vote, Millis timeout) {
  Writer w;
  w.i64(group_rank);
  w.i64(step);
  w.boolean(vote);
  std::string resp = rpc_.call(kManagerShouldCommit, w.data(), timeout);
  Reader r(resp);
  return r.boolean();
}

void ManagerClient::kill(const std::string& msg) {
  Writer w;
  w.str(msg);
  rpc_.call(kManagerKill, w.data(), Millis(10000));
}

}  // namespace tft
