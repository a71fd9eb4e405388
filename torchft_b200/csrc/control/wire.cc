#include "wire.h"

#include <pthread.h>

#include <arpa/inet.h>
#include <errno.h>
#include <fcntl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/types.h>
#include <unistd.h>

#include <algorithm>
#include <random>
#include <thread>

namespace tft {

void parse_addr(const std::string& addr_in, std::string* host, int* port) {
  std::string a = addr_in;
  const std::string schemes[] = {"http://", "https://", "tft://"};
  for (const auto& s : schemes)
    if (a.rfind(s, 0) == 0) a = a.substr(s.size());
  while (!a.empty() && a.back() == '/') a.pop_back();
  size_t colon = a.rfind(':');
  if (colon == std::string::npos) throw std::runtime_error("address has no port: " + addr_in);
  std::string h = a.substr(0, colon);
  if (!h.empty() && h.front() == '[' && h.back() == ']') h = h.substr(1, h.size() - 2);
  try {
    *port = std::stoi(a.substr(colon + 1));
  } catch (...) {
    throw std::runtime_error("bad port in address: " + addr_in);
  }
  *host = h;
}

static void set_common_opts(int fd) {
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
  setsockopt(fd, SOL_SOCKET, SO_KEEPALIVE, &one, sizeof(one));
#ifdef TCP_KEEPIDLE
  int idle = 60, intvl = 20, cnt = 3;  // reference: HTTP/2 keep-alive 60 s / 20 s (src/net.rs:19-24)
  setsockopt(fd, IPPROTO_TCP, TCP_KEEPIDLE, &idle, sizeof(idle));
  setsockopt(fd, IPPROTO_TCP, TCP_KEEPINTVL, &intvl, sizeof(intvl));
  setsockopt(fd, IPPROTO_TCP, TCP_KEEPCNT, &cnt, sizeof(cnt));
#endif
}

int listen_on(const std::string& bind_addr, int* port) {
  std::string host;
  int p = 0;
  parse_addr(bind_addr, &host, &p);
  int fd = -1;
  const bool v4 = !host.empty() && host.find(':') == std::string::npos && host != "::" && host != "*";
  if (!v4) {
    fd = ::socket(AF_INET6, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd >= 0) {
      int zero = 0, one = 1;
      setsockopt(fd, IPPROTO_IPV6, IPV6_V6ONLY, &zero, sizeof(zero));
      setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
      sockaddr_in6 sa{};
      sa.sin6_family = AF_INET6;
      sa.sin6_port = htons((uint16_t)p);
      if (host.empty() || host == "::" || host == "*") {
        sa.sin6_addr = in6addr_any;
      } else if (inet_pton(AF_INET6, host.c_str(), &sa.sin6_addr) != 1) {
        ::close(fd);
        fd = -1;
      }
      if (fd >= 0 && ::bind(fd, (sockaddr*)&sa, sizeof(sa)) != 0) {
        int e = errno;
        ::close(fd);
        fd = -1;
        if (e == EADDRINUSE) throw std::runtime_error("bind " + bind_addr + ": address in use");
      }
    }
  }
  if (fd < 0) {  // IPv4 (explicit v4 host, or no IPv6 stack in this container)
    fd = ::socket(AF_INET, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd < 0) throw std::runtime_error(std::string("socket: ") + strerror(errno));
    int one = 1;
    setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    sockaddr_in sa{};
    sa.sin_family = AF_INET;
    sa.sin_port = htons((uint16_t)p);
    if (!v4 || host == "0.0.0.0") {
      sa.sin_addr.s_addr = htonl(INADDR_ANY);
    } else if (inet_pton(AF_INET, host.c_str(), &sa.sin_addr) != 1) {
      ::close(fd);
      throw std::runtime_error("cannot parse bind host: " + host);
    }
    if (::bind(fd, (sockaddr*)&sa, sizeof(sa)) != 0) {
      std::string m = std::string("bind ") + bind_addr + ": " + strerror(errno);
      ::close(fd);
      throw std::runtime_error(m);
    }
  }
  if (::listen(fd, 1024) != 0) {
    std::string m = std::string("listen: ") + strerror(errno);
    ::close(fd);
    throw std::runtime_error(m);
  }
  sockaddr_storage ss{};
  socklen_t sl = sizeof(ss);
  getsockname(fd, (sockaddr*)&ss, &sl);
  *port = ss.ss_family == AF_INET6 ? ntohs(((sockaddr_in6*)&ss)->sin6_port) : ntohs(((sockaddr_in*)&ss)->sin_port);
  return fd;
}

static int try_connect_once(const std::string& host, int port, int timeout_ms, std::string* err) {
  addrinfo hints{}, *res = nullptr;
  hints.ai_family = AF_UNSPEC;
  hints.ai_socktype = SOCK_STREAM;
  std::string h = host;
  if (h.empty() || h == "::" || h == "0.0.0.0") h = "localhost";
  int rc = getaddrinfo(h.c_str(), std::to_string(port).c_str(), &hints, &res);
  if (rc != 0) {
    // container hostnames frequently do not resolve: fall back to loopback
    if (h == local_hostname()) rc = getaddrinfo("127.0.0.1", std::to_string(port).c_str(), &hints, &res);
    if (rc != 0) {
      *err = std::string("resolve ") + h + ": " + gai_strerror(rc);
      return -1;
    }
  }
  int fd = -1;
  for (addrinfo* ai = res; ai; ai = ai->ai_next) {
    fd = ::socket(ai->ai_family, SOCK_STREAM | SOCK_CLOEXEC | SOCK_NONBLOCK, 0);
    if (fd < 0) continue;
    int r = ::connect(fd, ai->ai_addr, ai->ai_addrlen);
    if (r != 0 && errno == EINPROGRESS) {
      pollfd pf{fd, POLLOUT, 0};
      r = ::poll(&pf, 1, std::max(1, timeout_ms));
      if (r == 1) {
        int so = 0;
        socklen_t sl = sizeof(so);
        getsockopt(fd, SOL_SOCKET, SO_ERROR, &so, &sl);
        r = so == 0 ? 0 : -1;
        if (so) errno = so;
      } else {
        r = -1;
        if (errno == 0) errno = ETIMEDOUT;
      }
    }
    if (r == 0) {
      int fl = fcntl(fd, F_GETFL, 0);
      fcntl(fd, F_SETFL, fl & ~O_NONBLOCK);
      set_common_opts(fd);
      break;
    }
    *err = std::string("connect ") + h + ":" + std::to_string(port) + ": " + strerror(errno);
    ::close(fd);
    fd = -1;
  }
  freeaddrinfo(res);
  return fd;
}

double Backoff::next() {
  static thread_local std::mt19937 rng{std::random_device{}()};
  if (current_ms <= 0)
    current_ms = initial_ms;
  else
    current_ms = std::min(max_ms, current_ms * factor);
  std::uniform_real_distribution<double> j(0.0, max_jitter_ms);
  return current_ms + j(rng);
}

int connect_with_backoff(const std::string& addr, TimePoint deadline) {
  std::string host;
  int port = 0;
  parse_addr(addr, &host, &port);
  Backoff bo;
  std::string err = "deadline already expired";
  while (true) {
    auto now = Clock::now();
    int left = (int)std::chrono::duration_cast<Millis>(deadline - now).count();
    if (left <= 0) throw TimeoutError("timed out connecting to " + addr + ": " + err);
    int fd = try_connect_once(host, port, std::min(left, 5000), &err);
    if (fd >= 0) return fd;
    double sl = bo.next();
    now = Clock::now();
    left = (int)std::chrono::duration_cast<Millis>(deadline - now).count();
    if (left <= 0) throw TimeoutError("timed out connecting to " + addr + ": " + err);
    std::this_thread::sleep_for(Millis((int)std::min<double>(sl, left)));
  }
}

static bool io_all(int fd, void* buf, size_t n, TimePoint deadline, bool write, bool* timed_out) {
  if (timed_out) *timed_out = false;
  char* p = (char*)buf;
  while (n > 0) {
    auto left = std::chrono::duration_cast<Millis>(deadline - Clock::now()).count();
    if (left <= 0) {
      if (timed_out) *timed_out = true;
      return false;
    }
    pollfd pf{fd, (short)(write ? POLLOUT : POLLIN), 0};
    int r = ::poll(&pf, 1, (int)std::min<long long>(left, 1000));
    if (r < 0) {
      if (errno == EINTR) continue;
      return false;
    }
    if (r == 0) continue;
    ssize_t k = write ? ::send(fd, p, n, MSG_NOSIGNAL) : ::recv(fd, p, n, 0);
    if (k < 0) {
      if (errno == EINTR || errno == EAGAIN) continue;
      return false;
    }
    if (k == 0) return false;  // EOF
    p += k;
    n -= (size_t)k;
  }
  return true;
}

bool send_all(int fd, const void* buf, size_t n, TimePoint deadline, bool* timed_out) {
  return io_all(fd, const_cast<void*>(buf), n, deadline, true, timed_out);
}
bool recv_all(int fd, void* buf, size_t n, TimePoint deadline, bool* timed_out) {
  return io_all(fd, buf, n, deadline, false, timed_out);
}

void name_this_thread(const std::string& name) {
  pthread_setname_np(pthread_self(), name.substr(0, 15).c_str());
}

void close_fd(int fd) {
  if (fd >= 0) ::close(fd);
}
void shutdown_fd(int fd) {
  if (fd >= 0) ::shutdown(fd, SHUT_RDWR);
}

std::string local_hostname() {
  char buf[256];
  if (gethostname(buf, sizeof(buf)) != 0) return "localhost";
  buf[sizeof(buf) - 1] = 0;
  return buf;
}

}  // namespace tft
