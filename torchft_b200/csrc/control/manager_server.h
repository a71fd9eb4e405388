// ManagerServer: one per replica group (hosted by group rank 0). Reference: src/manager.rs.
//   * heartbeats to the Lighthouse every `heartbeat_interval`
//   * Quorum RPC        = intra-group barrier; the last rank to arrive forwards ONE
//                         lighthouse quorum request (with retries) for the whole group
//   * ShouldCommit RPC  = intra-group AND-reduce barrier
//   * CheckpointMetadata: per-rank transport metadata served to healing peers
//   * Kill              : exit(1) (dashboard button / chaos tooling)
#pragma once

#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <thread>

#include "lighthouse.h"
#include "quorum.h"
#include "rpc.h"

namespace tft {

class ManagerServer : public RpcServer {
 public:
  ManagerServer(std::string replica_id, std::string lighthouse_addr, std::string hostname, const std::string& bind,
                std::string store_addr, uint64_t world_size, Millis heartbeat_interval, Millis connect_timeout,
                int64_t quorum_retries);
  ~ManagerServer() override;
  std::string address() const;
  void shutdown();

 protected:
  uint32_t handle_rpc(uint32_t method, const std::string& req, TimePoint deadline, std::string* resp) override;

 private:
  void heartbeat_loop();
  void run_quorum(QuorumMember requester, Millis timeout, uint64_t round);
  Quorum quorum_with_retries(const QuorumMember& requester, Millis timeout);
  std::shared_ptr<LighthouseClient> lighthouse_client(bool reconnect);

  const std::string replica_id_, lighthouse_addr_, hostname_, store_addr_;
  const uint64_t world_size_;
  const Millis heartbeat_interval_, connect_timeout_;
  const int64_t quorum_retries_;

  std::mutex mu_;
  std::condition_variable cv_;
  bool shutdown_ = false;
  std::map<int64_t, std::string> checkpoint_metadata_;
  std::map<int64_t, QuorumMember> participants_;
  // quorum broadcast: generation + (quorum | error)
  uint64_t quorum_gen_ = 0;
  Quorum latest_quorum_;
  std::string latest_error_;
  uint32_t latest_error_status_ = kStatusOk;
  // should_commit barrier
  std::set<int64_t> commit_count_, commit_failures_;
  uint64_t commit_gen_ = 0;
  bool commit_decision_ = false;

  std::mutex lh_mu_;
  std::shared_ptr<LighthouseClient> lh_client_;
  std::thread heartbeat_thread_;
  uint64_t active_round_ = 0;
  int quorum_workers_ = 0;
};

class ManagerClient {
 public:
  ManagerClient(const std::string& addr, Millis connect_timeout) : rpc_(addr, connect_timeout) {}
  QuorumResult quorum(int64_t group_rank, int64_t step, const std::string& checkpoint_metadata, bool shrink_only,
                      Millis timeout, int64_t commit_failures, bool init_sync);
  std::string checkpoint_metadata(int64_t rank, Millis timeout);
  bool should_commit(int64_t group_rank, int64_t step, bool should_commit, Millis timeout);
  void kill(const std::string& msg);

 private:
  RpcClient rpc_;
};

}  // namespace tft
