// Thread-per-connection RPC server/client over the framed protocol of wire.h (the reference gets this layer from
// tonic/hyper: src/net.rs:16-42 connect + keep-alive, src/timeout.rs:26-69 server-side deadlines).
// Thread-per-connection RPC/HTTP server base and the blocking RPC client.
#pragma once

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "wire.h"

namespace tft {

struct HttpRequest {
  std::string method;  // GET / POST
  std::string path;
};
struct HttpResponse {
  int code = 200;
  std::string content_type = "text/html; charset=utf-8";
  std::string body;
};

// One listening socket; every accepted connection gets a detached worker
// thread that serves framed RPCs (many per connection) or one HTTP request.
class RpcServer {
 public:
  RpcServer() = default;
  virtual ~RpcServer();

  void start(const std::string& bind, const std::string& thread_name);
  void stop();  // idempotent: closes listener + all live connections, waits for workers
  int port() const { return port_; }
  bool stopping() const { return stopping_.load(); }

 protected:
  // Return status + response payload. `deadline` is the client's deadline as
  // seen by the server; handlers must not block past it.
  virtual uint32_t handle_rpc(uint32_t method, const std::string& req, TimePoint deadline, std::string* resp) = 0;
  virtual HttpResponse handle_http(const HttpRequest& req);

 private:
  void accept_loop();
  void serve(int fd);

  int listen_fd_ = -1;
  int port_ = 0;
  std::thread accept_thread_;
  std::string thread_name_;
  std::atomic<bool> stopping_{false};
  std::mutex mu_;
  std::condition_variable cv_;
  std::set<int> conns_;
  int workers_ = 0;
};

// Blocking client with a small pool of persistent connections. Safe to call
// from several threads; concurrent calls use separate connections.
class RpcClient {
 public:
  RpcClient(std::string addr, Millis connect_timeout);
  ~RpcClient();
  // Throws TimeoutError on deadline (client- or server-side), RpcError otherwise.
  std::string call(uint32_t method, const std::string& payload, Millis timeout);
  const std::string& addr() const { return addr_; }
  // Abort every in-flight call (their sockets are shut down); used at shutdown.
  void cancel();

 private:
  int checkout(TimePoint deadline);
  void checkin(int fd);

  std::string addr_;
  Millis connect_timeout_;
  std::mutex mu_;
  std::vector<int> idle_;
  std::set<int> busy_;
};

}  // namespace tft
