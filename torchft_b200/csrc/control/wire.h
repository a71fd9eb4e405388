// Framed binary RPC + socket helpers for the torchft_b200 control plane.
//
// The reference speaks gRPC/HTTP2 via tonic (src/net.rs, src/timeout.rs,
// proto/torchft.proto). Neither gRPC C++ nor protoc exists in this image, so
// the control plane uses a small length-prefixed protocol with the same
// message fields and the same semantics that matter:
//   * every request carries the client's timeout and the SERVER honours it
//     (reference: src/timeout.rs:26-69 parses the grpc-timeout header);
//   * deadline/cancel map to Python TimeoutError, everything else to
//     RuntimeError (reference: src/lib.rs:673-697);
//   * connect uses exponential backoff (100 ms -> 10 s, x1.5, <=100 ms jitter)
//     under an overall deadline (reference: src/retry.rs:14-49, src/net.rs:16-42);
//   * the same listening port also answers plain HTTP/1.1 (dashboard).
//
// Request : "TFT1" u32 method | u64 timeout_ms | u32 len | payload
// Response: u32 status | u32 len | payload (or UTF-8 error message)
#pragma once

#include <chrono>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace tft {

using Clock = std::chrono::steady_clock;
using TimePoint = Clock::time_point;
using Millis = std::chrono::milliseconds;

enum Status : uint32_t {
  kStatusOk = 0,
  kStatusDeadline = 1,
  kStatusCancelled = 2,
  kStatusInvalid = 3,
  kStatusNotFound = 4,
  kStatusInternal = 5,
  kStatusUnavailable = 6,
};

enum Method : uint32_t {
  kLighthouseQuorum = 1,
  kLighthouseHeartbeat = 2,
  kManagerQuorum = 10,
  kManagerCheckpointMetadata = 11,
  kManagerShouldCommit = 12,
  kManagerKill = 13,
};

constexpr char kMagic[4] = {'T', 'F', 'T', '1'};

struct TimeoutError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct RpcError : std::runtime_error {
  uint32_t status;
  RpcError(uint32_t s, const std::string& m) : std::runtime_error(m), status(s) {}
};

// ----------------------------------------------------------------- encoding
class Writer {
 public:
  void u8(uint8_t v) { buf_.push_back((char)v); }
  void u32(uint32_t v) { raw(&v, 4); }
  void u64(uint64_t v) { raw(&v, 8); }
  void i64(int64_t v) { raw(&v, 8); }
  void boolean(bool v) { u8(v ? 1 : 0); }
  void str(const std::string& s) {
    u32((uint32_t)s.size());
    buf_.append(s);
  }
  void raw(const void* p, size_t n) { buf_.append((const char*)p, n); }
  const std::string& data() const { return buf_; }
  std::string take() { return std::move(buf_); }

 private:
  std::string buf_;
};

class Reader {
 public:
  explicit Reader(const std::string& b) : p_(b.data()), end_(b.data() + b.size()) {}
  uint8_t u8() {
    need(1);
    return (uint8_t)*p_++;
  }
  uint32_t u32() {
    uint32_t v;
    get(&v, 4);
    return v;
  }
  uint64_t u64() {
    uint64_t v;
    get(&v, 8);
    return v;
  }
  int64_t i64() {
    int64_t v;
    get(&v, 8);
    return v;
  }
  bool boolean() { return u8() != 0; }
  std::string str() {
    uint32_t n = u32();
    need(n);
    std::string s(p_, n);
    p_ += n;
    return s;
  }
  bool done() const { return p_ == end_; }

 private:
  void need(size_t n) {
    if ((size_t)(end_ - p_) < n) throw RpcError(kStatusInvalid, "truncated message");
  }
  void get(void* out, size_t n) {
    need(n);
    std::memcpy(out, p_, n);
    p_ += n;
  }
  const char* p_;
  const char* end_;
};

// ------------------------------------------------------------------ sockets
// "http://host:port", "host:port", "[::]:port" -> (host, port)
void parse_addr(const std::string& addr, std::string* host, int* port);

// Listen on `bind` ("[::]:0", "0.0.0.0:1234", ...). Returns fd; *port gets the bound port.
int listen_on(const std::string& bind, int* port);

// Connect with exponential backoff until `deadline`. Throws TimeoutError.
int connect_with_backoff(const std::string& addr, TimePoint deadline);

// Blocking exact-size IO bounded by `deadline`; false on EOF/error/timeout
// (*timed_out tells which).
bool send_all(int fd, const void* buf, size_t n, TimePoint deadline, bool* timed_out = nullptr);
bool recv_all(int fd, void* buf, size_t n, TimePoint deadline, bool* timed_out = nullptr);

// Name the calling thread (<= 15 chars shown by top -H / gdb / py-spy); reference: tokio thread names, src/lib.rs:105,166,500,635.
void name_this_thread(const std::string& name);
void close_fd(int fd);
void shutdown_fd(int fd);
std::string local_hostname();

// Exponential backoff schedule (exposed for unit tests).
struct Backoff {
  double initial_ms = 100, max_ms = 10000, factor = 1.5, max_jitter_ms = 100;
  double current_ms = 0;
  // next sleep in ms (with jitter in [0, max_jitter_ms])
  double next();
};

}  // namespace tft
