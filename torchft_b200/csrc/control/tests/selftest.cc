// Native unit + integration tests of the C++ control plane (the role of the reference's
// `cargo test`: src/lighthouse.rs:627-1296, src/manager.rs:656-1218, retry.rs, timeout.rs).
// No test framework dependency: CHECK() records failures, main() returns their count.
//
//   build:  python -m torchft_b200._build   (-> bin/torchft_b200_selftest)
//   run:    bin/torchft_b200_selftest [filter]
#include <arpa/inet.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <functional>
#include <future>
#include <string>
#include <thread>
#include <vector>

#include "../lighthouse.h"
#include "../manager_server.h"
#include "../quorum.h"
#include "../wire.h"

using namespace tft;

static int g_failed = 0, g_checks = 0;
#define CHECK(cond)                                                              \
  do {                                                                           \
    ++g_checks;                                                                  \
    if (!(cond)) {                                                               \
      ++g_failed;                                                                \
      std::fprintf(stderr, "  FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);   \
    }                                                                            \
  } while (0)

static QuorumMember member(const std::string& id, int64_t step = 1, bool shrink_only = false) {
  QuorumMember m;
  m.replica_id = id;
  m.address = "http://addr_" + id;
  m.store_address = "store_" + id + ":1";
  m.step = step;
  m.world_size = 1;
  m.shrink_only = shrink_only;
  return m;
}

static void join(LighthouseState& s, const std::string& id, int64_t now, int64_t step = 1, bool shrink_only = false) {
  s.participants[id] = ParticipantDetails{now, member(id, step, shrink_only)};
  s.heartbeats[id] = now;
}

// ---------------------------------------------------------------- pure logic
static void test_wire_roundtrip() {
  Quorum q;
  q.quorum_id = 42;
  q.created_ms = 123456789;
  q.participants = {member("a", 3), member("b", 7, true)};
  q.participants[1].data = "{\"k\": [1, 2]}";
  q.participants[1].commit_failures = 5;
  Writer w;
  q.encode(w);
  const std::string bytes = w.take();
  Reader r(bytes);
  Quorum d = Quorum::decode(r);
  CHECK(d.quorum_id == 42 && d.created_ms == 123456789 && d.participants.size() == 2);
  CHECK(d.participants[1].replica_id == "b" && d.participants[1].shrink_only && d.participants[1].step == 7);
  CHECK(d.participants[1].data == "{\"k\": [1, 2]}" && d.participants[1].commit_failures == 5);
  // truncated input must throw, never read out of bounds
  bool threw = false;
  try {
    const std::string cut = bytes.substr(0, bytes.size() / 2);
    Reader r2(cut);
    Quorum::decode(r2);
  } catch (const RpcError&) {
    threw = true;
  }
  CHECK(threw);
}

static void test_backoff_schedule() {
  Backoff b;
  b.max_jitter_ms = 0;
  double prev = 0;
  for (int i = 0; i < 40; ++i) {
    const double d = b.next();
    CHECK(d >= prev || d == b.max_ms);
    CHECK(d <= b.max_ms);
    prev = d;
  }
  CHECK(prev == b.max_ms);
  Backoff j;
  for (int i = 0; i < 10; ++i) {
    const double d = j.next();
    CHECK(d >= 100.0 && d <= j.max_ms + j.max_jitter_ms);
  }
}

static void test_quorum_join_timeout() {
  LighthouseOpt opt;
  opt.min_replicas = 1;
  opt.join_timeout_ms = 60000;
  LighthouseState s;
  const int64_t now = 1000000;
  CHECK(!quorum_compute(now, s, opt).participants.has_value());  // nobody
  join(s, "a", now);
  // first quorum, nobody else known: valid immediately once every healthy replica has joined
  CHECK(quorum_compute(now, s, opt).participants.has_value());
  // a third replica heart-beats but has not joined: a majority (2 of 3) is here, so the quorum is
  // valid but held back for the straggler until the join timeout
  join(s, "b", now);
  s.heartbeats["c"] = now;
  CHECK(!quorum_compute(now, s, opt).participants.has_value());
  const int64_t later = now + opt.join_timeout_ms + 1;
  for (const char* id : {"a", "b", "c"}) s.heartbeats[id] = later;  // all still alive when it expires
  const auto d = quorum_compute(later, s, opt);
  CHECK(d.participants.has_value() && d.participants->size() == 2);
}

static void test_quorum_heartbeat_expiry_and_min_replicas() {
  LighthouseOpt opt;
  opt.min_replicas = 2;
  opt.join_timeout_ms = 0;
  opt.heartbeat_timeout_ms = 5000;
  LighthouseState s;
  const int64_t now = 1000000;
  join(s, "a", now);
  CHECK(!quorum_compute(now, s, opt).participants.has_value());  // below min_replicas
  join(s, "b", now);
  auto d = quorum_compute(now, s, opt);
  CHECK(d.participants.has_value() && d.participants->size() == 2);
  // b stops heart-beating: it no longer counts
  s.heartbeats["b"] = now - 6000;
  CHECK(!quorum_compute(now, s, opt).participants.has_value());
}

static void test_quorum_fast_path_and_shrink_only() {
  LighthouseOpt opt;
  opt.min_replicas = 1;
  opt.join_timeout_ms = 60000;
  LighthouseState s;
  const int64_t now = 1000000;
  Quorum prev;
  prev.quorum_id = 1;
  prev.participants = {member("a"), member("b")};
  s.prev_quorum = prev;
  s.quorum_id = 1;
  join(s, "a", now);
  s.heartbeats["b"] = now;
  CHECK(!quorum_compute(now, s, opt).participants.has_value());  // b healthy but not here yet
  join(s, "b", now);
  auto d = quorum_compute(now, s, opt);  // all of the previous quorum is back: no waiting for stragglers
  CHECK(d.participants.has_value() && d.participants->size() == 2);
  // a third replica wants in while someone asks shrink_only: it is left out
  join(s, "c", now);
  s.participants["a"].member.shrink_only = true;
  d = quorum_compute(now, s, opt);
  CHECK(d.participants.has_value() && d.participants->size() == 2);
  for (const auto& m : *d.participants) CHECK(m.replica_id != "c");
}

static void test_quorum_split_brain_guard() {
  LighthouseOpt opt;
  opt.min_replicas = 1;
  opt.join_timeout_ms = 0;
  LighthouseState s;
  const int64_t now = 1000000;
  join(s, "a", now);
  for (const char* id : {"b", "c"}) s.heartbeats[id] = now;  // healthy, not participating
  // 1 of 3 healthy replicas is not a majority: refuse (two halves could otherwise both proceed)
  CHECK(!quorum_compute(now, s, opt).participants.has_value());
  join(s, "b", now);
  CHECK(quorum_compute(now, s, opt).participants.has_value());
}

static void test_quorum_changed() {
  std::vector<QuorumMember> a = {member("a"), member("b")}, b = {member("a"), member("b")};
  CHECK(!quorum_changed(a, b));
  b[1].step = 99;  // only membership matters
  CHECK(!quorum_changed(a, b));
  b.push_back(member("c"));
  CHECK(quorum_changed(a, b));
}

static void test_compute_quorum_results() {
  Quorum q;
  q.quorum_id = 7;
  q.participants = {member("r0", 10), member("r1", 10), member("r2", 8), member("r3", 0)};
  // up-to-date replica: not healing, serves the stragglers assigned to it
  QuorumResult r0 = compute_quorum_results("r0", 0, q, true);
  CHECK(r0.quorum_id == 7 && r0.replica_rank == 0 && r0.replica_world_size == 4);
  CHECK(!r0.heal && r0.max_step == 10 && r0.max_world_size == 2);
  QuorumResult r1 = compute_quorum_results("r1", 0, q, true);
  std::vector<int64_t> served = r0.recover_dst_replica_ranks;
  served.insert(served.end(), r1.recover_dst_replica_ranks.begin(), r1.recover_dst_replica_ranks.end());
  CHECK(served.size() == 2);  // r2 and r3 are spread over the two up-to-date replicas
  QuorumResult r2 = compute_quorum_results("r2", 0, q, true);
  CHECK(r2.heal && r2.recover_src_replica_rank.has_value());
  CHECK(*r2.recover_src_replica_rank == 0 || *r2.recover_src_replica_rank == 1);
  CHECK(!r2.recover_src_manager_address.empty());
  // every rank agrees on the primary store
  CHECK(r0.store_address == r2.store_address && !r0.store_address.empty());
  // unknown replica -> NotFound
  bool threw = false;
  try {
    compute_quorum_results("nope", 0, q, true);
  } catch (const RpcError&) {
    threw = true;
  }
  CHECK(threw);
  // first step with init_sync=false: nobody heals even though steps are all 0
  Quorum q0;
  q0.quorum_id = 1;
  q0.participants = {member("x", 0), member("y", 0)};
  CHECK(!compute_quorum_results("y", 0, q0, false).heal);
  // ... with init_sync=true everybody but the primary syncs from it
  CHECK(compute_quorum_results("y", 0, q0, true).heal != compute_quorum_results("x", 0, q0, true).heal);
}

// ------------------------------------------------------------ integration
static void test_lighthouse_end_to_end() {
  LighthouseOpt opt;
  opt.bind = "127.0.0.1:0";
  opt.min_replicas = 2;
  opt.join_timeout_ms = 100;
  opt.quorum_tick_ms = 10;
  Lighthouse lh(opt);
  const std::string addr = lh.address();
  auto ask = [&](const std::string& id, int64_t step) {
    LighthouseClient c(addr, Millis(5000));
    return c.quorum(member(id, step), Millis(10000));
  };
  auto fa = std::async(std::launch::async, ask, "a", 1);
  auto fb = std::async(std::launch::async, ask, "b", 1);
  Quorum qa = fa.get(), qb = fb.get();
  CHECK(qa.quorum_id == qb.quorum_id && qa.participants.size() == 2);
  // a client that is alone never gets a quorum: the deadline surfaces as TimeoutError quickly
  LighthouseClient lonely(addr, Millis(5000));
  bool timed_out = false;
  const auto t0 = std::chrono::steady_clock::now();
  try {
    lonely.quorum(member("c", 1), Millis(300));
  } catch (const TimeoutError&) {
    timed_out = true;
  }
  const auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
  CHECK(timed_out && ms < 3000);
  lonely.heartbeat("c", Millis(1000));
  lh.shutdown();
}

static void test_manager_group_barrier_and_commit() {
  LighthouseOpt opt;
  opt.bind = "127.0.0.1:0";
  opt.min_replicas = 1;
  opt.join_timeout_ms = 100;
  opt.quorum_tick_ms = 10;
  Lighthouse lh(opt);
  // one replica group with TWO ranks: quorum() and should_commit() are group-wide barriers
  ManagerServer ms("group0", lh.address(), "127.0.0.1", "127.0.0.1:0", "store:1", 2, Millis(50), Millis(5000), 0);
  auto rank_fn = [&](int64_t rank, bool vote) {
    ManagerClient c(ms.address(), Millis(5000));
    QuorumResult q = c.quorum(rank, 5, "meta" + std::to_string(rank), false, Millis(10000), 0, true);
    const bool ok = c.should_commit(rank, 5, vote, Millis(10000));
    return std::make_pair(q, ok);
  };
  auto f0 = std::async(std::launch::async, rank_fn, 0, true);
  auto f1 = std::async(std::launch::async, rank_fn, 1, true);
  auto r0 = f0.get(), r1 = f1.get();
  CHECK(r0.first.quorum_id == r1.first.quorum_id && r0.first.replica_world_size == 1);
  CHECK(r0.second && r1.second);
  // one "no" vote vetoes the step for the whole group
  auto g0 = std::async(std::launch::async, [&] { return ManagerClient(ms.address(), Millis(5000)).should_commit(0, 6, true, Millis(10000)); });
  auto g1 = std::async(std::launch::async, [&] { return ManagerClient(ms.address(), Millis(5000)).should_commit(1, 6, false, Millis(10000)); });
  CHECK(!g0.get());
  CHECK(!g1.get());
  // metadata published with the quorum request is served to healing peers
  ManagerClient c(ms.address(), Millis(5000));
  CHECK(c.checkpoint_metadata(1, Millis(5000)) == "meta1");
  // a barrier nobody else joins times out instead of hanging
  bool timed_out = false;
  try {
    c.should_commit(0, 7, true, Millis(200));
  } catch (const TimeoutError&) {
    timed_out = true;
  }
  CHECK(timed_out);
  ms.shutdown();
  lh.shutdown();
}

// Raw-socket abuse of a live server; most valuable under ASAN/UBSAN (scripts/sanitize.sh).
static int raw_connect(const std::string& http_addr) {
  const auto colon = http_addr.rfind(':');
  const int port = std::stoi(http_addr.substr(colon + 1));
  int fd = ::socket(AF_INET, SOCK_STREAM, 0);
  sockaddr_in sa{};
  sa.sin_family = AF_INET;
  sa.sin_port = htons((uint16_t)port);
  sa.sin_addr.s_addr = htonl(INADDR_LOOPBACK);
  if (::connect(fd, (sockaddr*)&sa, sizeof(sa)) != 0) {
    ::close(fd);
    return -1;
  }
  return fd;
}

static void test_server_survives_garbage() {
  LighthouseOpt opt;
  opt.bind = "127.0.0.1:0";
  opt.min_replicas = 1;
  opt.join_timeout_ms = 50;
  opt.quorum_tick_ms = 10;
  Lighthouse lh(opt);
  uint64_t x = 0x9e3779b97f4a7c15ull;  // xorshift: deterministic "random" bytes
  auto next = [&x] {
    x ^= x << 13;
    x ^= x >> 7;
    x ^= x << 17;
    return x;
  };
  for (int round = 0; round < 60; ++round) {
    int fd = raw_connect(lh.address());
    CHECK(fd >= 0);
    if (fd < 0) return;
    std::string msg;
    const int kind = round % 5;
    if (kind == 0) {  // pure noise
      for (int i = 0; i < 200; ++i) msg.push_back((char)next());
    } else {          // valid magic, then a header with hostile fields and a short/garbage payload
      msg = "TFT1";
      uint32_t method = kind == 1 ? 1u : (kind == 2 ? 2u : (uint32_t)next());
      uint64_t timeout_ms = kind == 3 ? ~0ull : 50;
      uint32_t len = kind == 4 ? 0x7fffffffu : (uint32_t)(next() % 64);
      msg.append((const char*)&method, 4);
      msg.append((const char*)&timeout_ms, 8);
      msg.append((const char*)&len, 4);
      const int n = (int)(next() % 48);
      for (int i = 0; i < n; ++i) msg.push_back((char)next());
    }
    (void)!::send(fd, msg.data(), msg.size(), MSG_NOSIGNAL);
    if (round % 2) ::shutdown(fd, SHUT_WR);
    char buf[256];
    timeval tv{0, 20000};
    ::setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
    (void)!::recv(fd, buf, sizeof(buf), 0);
    ::close(fd);
  }
  // still fully functional
  LighthouseClient c(lh.address(), Millis(5000));
  Quorum q = c.quorum(member("after_garbage", 1), Millis(5000));
  CHECK(q.participants.size() == 1 && q.participants[0].replica_id == "after_garbage");
  lh.shutdown();
}

int main(int argc, char** argv) {
  const char* filter = argc > 1 ? argv[1] : "";
  struct T {
    const char* name;
    std::function<void()> fn;
  };
  const std::vector<T> tests = {
      {"wire_roundtrip", test_wire_roundtrip},
      {"backoff_schedule", test_backoff_schedule},
      {"quorum_join_timeout", test_quorum_join_timeout},
      {"quorum_heartbeat_expiry_and_min_replicas", test_quorum_heartbeat_expiry_and_min_replicas},
      {"quorum_fast_path_and_shrink_only", test_quorum_fast_path_and_shrink_only},
      {"quorum_split_brain_guard", test_quorum_split_brain_guard},
      {"quorum_changed", test_quorum_changed},
      {"compute_quorum_results", test_compute_quorum_results},
      {"lighthouse_end_to_end", test_lighthouse_end_to_end},
      {"manager_group_barrier_and_commit", test_manager_group_barrier_and_commit},
      {"server_survives_garbage", test_server_survives_garbage},
  };
  int ran = 0;
  for (const auto& t : tests) {
    if (*filter && !std::strstr(t.name, filter)) continue;
    const int before = g_failed;
    std::fprintf(stderr, "[ RUN  ] %s\n", t.name);
    try {
      t.fn();
    } catch (const std::exception& e) {
      ++g_failed;
      std::fprintf(stderr, "  EXCEPTION: %s\n", e.what());
    }
    std::fprintf(stderr, "[ %s ] %s\n", g_failed == before ? " OK " : "FAIL", t.name);
    ++ran;
  }
  std::printf("selftest: %d tests, %d checks, %d failed\n", ran, g_checks, g_failed);
  return g_failed ? 1 : 0;
}
