// Lighthouse: global quorum authority (reference: src/lighthouse.rs).
#pragma once

#include <condition_variable>
#include <deque>
#include <memory>
#include <set>
#include <mutex>
#include <thread>

#include "quorum.h"
#include "rpc.h"

namespace tft {

class Lighthouse : public RpcServer {
 public:
  explicit Lighthouse(LighthouseOpt opt);
  ~Lighthouse() override;
  std::string address() const;  // "http://<hostname>:<port>"
  void shutdown();

 protected:
  uint32_t handle_rpc(uint32_t method, const std::string& req, TimePoint deadline, std::string* resp) override;
  HttpResponse handle_http(const HttpRequest& req) override;

 private:
  void tick_locked();  // requires mu_
  void tick_loop();
  std::string status_html();
  std::string status_json();
  HttpResponse kill_replica(const std::string& replica_id);

  LighthouseOpt opt_;
  std::mutex mu_;
  std::condition_variable cv_;
  LighthouseState state_;
  // every formed quorum gets a generation number; waiters replay the ones they missed
  uint64_t gen_ = 0;
  // (generation, quorum, its wire encoding, member ids): encoded ONCE when the quorum forms; every waiter of a large
  // quorum used to re-encode all N members under the lock (O(N^2) bytes per round)
  struct Formed {
    uint64_t gen;
    Quorum quorum;
    std::shared_ptr<const std::string> wire;
    std::set<std::string> members;
  };
  std::deque<Formed> history_;
  std::string last_reason_;
  std::thread tick_thread_;
  bool shutdown_ = false;
};

// ---- clients ---------------------------------------------------------------
class LighthouseClient {
 public:
  LighthouseClient(const std::string& addr, Millis connect_timeout) : rpc_(addr, connect_timeout) {}
  Quorum quorum(const QuorumMember& requester, Millis timeout);
  void heartbeat(const std::string& replica_id, Millis timeout);
  void cancel() { rpc_.cancel(); }

 private:
  RpcClient rpc_;
};

}  // namespace tft
