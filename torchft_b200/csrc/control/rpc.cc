#include "rpc.h"

#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <sstream>

namespace tft {

RpcServer::~RpcServer() { stop(); }

void RpcServer::start(const std::string& bind, const std::string& thread_name) {
  listen_fd_ = listen_on(bind, &port_);
  thread_name_ = thread_name;
  accept_thread_ = std::thread([this] {
    name_this_thread(thread_name_ + "-acc");
    accept_loop();
  });
}

void RpcServer::stop() {
  bool was = stopping_.exchange(true);
  if (was && !accept_thread_.joinable()) return;
  shutdown_fd(listen_fd_);
  {
    std::lock_guard<std::mutex> g(mu_);
    for (int fd : conns_) shutdown_fd(fd);
  }
  if (accept_thread_.joinable()) accept_thread_.join();
  close_fd(listen_fd_);
  listen_fd_ = -1;
  std::unique_lock<std::mutex> lk(mu_);
  cv_.wait_for(lk, std::chrono::seconds(5), [this] { return workers_ == 0; });
}

void RpcServer::accept_loop() {
  while (!stopping_.load()) {
    pollfd pf{listen_fd_, POLLIN, 0};
    int r = ::poll(&pf, 1, 200);
    if (r <= 0) continue;
    int fd = ::accept4(listen_fd_, nullptr, nullptr, SOCK_CLOEXEC);
    if (fd < 0) continue;
    int one = 1;
    setsockopt(fd, 6 /*IPPROTO_TCP*/, 1 /*TCP_NODELAY*/, &one, sizeof(one));
    {
      std::lock_guard<std::mutex> g(mu_);
      if (stopping_.load()) {
        close_fd(fd);
        break;
      }
      conns_.insert(fd);
      ++workers_;
    }
    std::thread([this, fd] {
      name_this_thread(thread_name_ + "-conn");
      try {
        serve(fd);
      } catch (...) {
      }
      close_fd(fd);
      {
        // Last touch of `this`: stop() may return (and the server be destroyed) the moment it observes
        // workers_ == 0, so the notify must happen INSIDE the critical section, never after it
        // (found by ThreadSanitizer: cv_ destroyed under a late notify_all).
        std::lock_guard<std::mutex> g(mu_);
        conns_.erase(fd);
        --workers_;
        cv_.notify_all();
      }
    }).detach();
  }
}

HttpResponse RpcServer::handle_http(const HttpRequest&) { return {404, "text/plain", "not found"}; }

static bool read_http_head(int fd, const char first[4], std::string* head) {
  head->assign(first, 4);
  char c;
  auto deadline = Clock::now() + std::chrono::seconds(10);
  while (head->size() < 16384) {
    if (head->size() >= 4 && head->compare(head->size() - 4, 4, "\r\n\r\n") == 0) return true;
    if (!recv_all(fd, &c, 1, deadline)) return false;
    head->push_back(c);
  }
  return false;
}

void RpcServer::serve(int fd) {
  while (!stopping_.load()) {
    char magic[4];
    // idle connections may sit here for a long time; poll in slices so stop() is prompt
    bool got = false;
    while (!stopping_.load()) {
      pollfd pf{fd, POLLIN, 0};
      int r = ::poll(&pf, 1, 500);
      if (r < 0) return;
      if (r == 0) continue;
      got = recv_all(fd, magic, 4, Clock::now() + std::chrono::seconds(30));
      break;
    }
    if (!got) return;
    if (std::memcmp(magic, kMagic, 4) != 0) {
      // plain HTTP/1.1 on the same port (dashboard, kill button)
      std::string head;
      if (!read_http_head(fd, magic, &head)) return;
      HttpRequest req;
      std::istringstream is(head);
      is >> req.method >> req.path;
      HttpResponse resp = handle_http(req);
      std::ostringstream os;
      const char* reason = resp.code == 200 ? "OK" : (resp.code == 404 ? "Not Found" : "Error");
      os << "HTTP/1.1 " << resp.code << " " << reason << "\r\nContent-Type: " << resp.content_type
         << "\r\nContent-Length: " << resp.body.size() << "\r\nConnection: close\r\n\r\n"
         << resp.body;
      std::string out = os.str();
      send_all(fd, out.data(), out.size(), Clock::now() + std::chrono::seconds(10));
      return;
    }
    struct {
      uint32_t method;
      uint64_t timeout_ms;
      uint32_t len;
    } __attribute__((packed)) hdr;
    auto io_deadline = Clock::now() + std::chrono::seconds(30);
    if (!recv_all(fd, &hdr, sizeof(hdr), io_deadline)) return;
    if (hdr.len > (64u << 20)) return;
    std::string payload(hdr.len, '\0');
    if (hdr.len && !recv_all(fd, payload.data(), hdr.len, io_deadline)) return;
    const TimePoint deadline = Clock::now() + Millis(std::min<uint64_t>(hdr.timeout_ms, 1000ull * 3600 * 24 * 30));
    std::string resp;
    uint32_t status;
    try {
      status = handle_rpc(hdr.method, payload, deadline, &resp);
    } catch (const RpcError& e) {
      status = e.status;
      resp = e.what();
    } catch (const TimeoutError& e) {
      status = kStatusDeadline;
      resp = e.what();
    } catch (const std::exception& e) {
      status = kStatusInternal;
      resp = e.what();
    }
    struct {
      uint32_t status;
      uint32_t len;
    } __attribute__((packed)) rh{status, (uint32_t)resp.size()};
    auto wd = Clock::now() + std::chrono::seconds(30);
    if (!send_all(fd, &rh, sizeof(rh), wd)) return;
    if (!resp.empty() && !send_all(fd, resp.data(), resp.size(), wd)) return;
  }
}

// ------------------------------------------------------------------- client
RpcClient::RpcClient(std::string addr, Millis connect_timeout)
    : addr_(std::move(addr)), connect_timeout_(connect_timeout) {
  // Fail fast (like the reference's eager channel connect): establish one connection now.
  int fd = connect_with_backoff(addr_, Clock::now() + connect_timeout_);
  idle_.push_back(fd);
}

RpcClient::~RpcClient() {
  std::lock_guard<std::mutex> g(mu_);
  for (int fd : idle_) close_fd(fd);
  idle_.clear();
}

int RpcClient::checkout(TimePoint deadline) {
  {
    std::lock_guard<std::mutex> g(mu_);
    while (!idle_.empty()) {
      int fd = idle_.back();
      idle_.pop_back();
      // drop connections the peer already closed
      pollfd pf{fd, POLLIN, 0};
      if (::poll(&pf, 1, 0) == 0) return fd;
      close_fd(fd);
    }
  }
  auto cd = std::min(deadline, Clock::now() + connect_timeout_);
  return connect_with_backoff(addr_, cd);
}

void RpcClient::cancel() {
  std::lock_guard<std::mutex> g(mu_);
  for (int fd : busy_) shutdown_fd(fd);
}

void RpcClient::checkin(int fd) {
  std::lock_guard<std::mutex> g(mu_);
  busy_.erase(fd);
  if (idle_.size() < 4)
    idle_.push_back(fd);
  else
    close_fd(fd);
}

std::string RpcClient::call(uint32_t method, const std::string& payload, Millis timeout) {
  const TimePoint deadline = Clock::now() + timeout;
  int fd = checkout(deadline);
  {
    std::lock_guard<std::mutex> g(mu_);
    busy_.insert(fd);
  }
  auto drop = [this](int f) {
    {
      std::lock_guard<std::mutex> g(mu_);
      busy_.erase(f);
    }
    close_fd(f);
  };
  struct {
    char magic[4];
    uint32_t method;
    uint64_t timeout_ms;
    uint32_t len;
  } __attribute__((packed)) hdr;
  std::memcpy(hdr.magic, kMagic, 4);
  hdr.method = method;
  hdr.timeout_ms = (uint64_t)std::max<int64_t>(0, timeout.count());
  hdr.len = (uint32_t)payload.size();
  bool to = false;
  // the server answers DEADLINE itself at `deadline`; give the reply a grace
  // period to arrive before declaring a client-side timeout
  const TimePoint io_deadline = deadline + Millis(1000);
  if (!send_all(fd, &hdr, sizeof(hdr), io_deadline, &to) ||
      (!payload.empty() && !send_all(fd, payload.data(), payload.size(), io_deadline, &to))) {
    drop(fd);
    if (to) throw TimeoutError("rpc to " + addr_ + " timed out while sending");
    throw RpcError(kStatusUnavailable, "connection to " + addr_ + " lost while sending request");
  }
  struct {
    uint32_t status;
    uint32_t len;
  } __attribute__((packed)) rh;
  if (!recv_all(fd, &rh, sizeof(rh), io_deadline, &to)) {
    drop(fd);
    if (to) throw TimeoutError("rpc to " + addr_ + " timed out after " + std::to_string(timeout.count()) + " ms");
    throw RpcError(kStatusUnavailable, "connection to " + addr_ + " closed before a response arrived");
  }
  std::string resp(rh.len, '\0');
  if (rh.len && !recv_all(fd, resp.data(), rh.len, io_deadline + Millis(5000), &to)) {
    drop(fd);
    throw RpcError(kStatusUnavailable, "connection to " + addr_ + " lost mid-response");
  }
  checkin(fd);
  if (rh.status == kStatusOk) return resp;
  if (rh.status == kStatusDeadline || rh.status == kStatusCancelled) throw TimeoutError(resp);
  throw RpcError(rh.status, resp);
}

}  // namespace tft
