// pybind11 module `torchft_b200._C`: the control-plane classes the Python
// Manager drives. Counterpart of the reference's pyo3 module (src/lib.rs):
// every blocking call releases the GIL; deadline/cancel surface as Python
// TimeoutError and everything else as RuntimeError (src/lib.rs:673-697).
#include <pybind11/chrono.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <csignal>
#include <cstdio>
#include <iostream>

#include "lighthouse.h"
#include "manager_server.h"
#include "quorum.h"

namespace py = pybind11;
using namespace tft;

namespace {

struct Timestamp {
  int64_t seconds = 0;
  int32_t nanos = 0;
};

Millis to_ms(std::chrono::duration<double> d) {
  return Millis((int64_t)(d.count() * 1000.0 + 0.5));
}

py::object json_loads(const std::string& s) {
  if (s.empty()) return py::none();
  return py::module_::import("json").attr("loads")(s);
}
std::string json_dumps(const py::object& o) {
  if (o.is_none()) return "";
  return py::cast<std::string>(py::module_::import("json").attr("dumps")(o));
}

int lighthouse_main_impl(const std::vector<std::string>& argv);

}  // namespace

int run_lighthouse_cli(const std::vector<std::string>& args);  // lighthouse_cli.cc

PYBIND11_MODULE(_C, m) {
  m.doc() = "torchft_b200 control plane: Lighthouse, ManagerServer and clients (C++17)";

  py::register_exception_translator([](std::exception_ptr p) {
    try {
      if (p) std::rethrow_exception(p);
    } catch (const TimeoutError& e) {
      PyErr_SetString(PyExc_TimeoutError, e.what());
    } catch (const RpcError& e) {
      PyErr_SetString(PyExc_RuntimeError, e.what());
    }
  });

  py::class_<Timestamp>(m, "Timestamp")
      .def(py::init<>())
      .def_readwrite("seconds", &Timestamp::seconds)
      .def_readwrite("nanos", &Timestamp::nanos);

  py::class_<QuorumMember>(m, "QuorumMember")
      .def(py::init([](std::string replica_id, std::string address, std::string store_address, int64_t step,
                       uint64_t world_size, bool shrink_only, py::object data, int64_t commit_failures) {
             QuorumMember q;
             q.replica_id = std::move(replica_id);
             q.address = std::move(address);
             q.store_address = std::move(store_address);
             q.step = step;
             q.world_size = world_size;
             q.shrink_only = shrink_only;
             q.data = json_dumps(data);
             q.commit_failures = commit_failures;
             return q;
           }),
           py::arg("replica_id") = "", py::arg("address") = "", py::arg("store_address") = "", py::arg("step") = 0,
           py::arg("world_size") = 0, py::arg("shrink_only") = false, py::arg("data") = py::none(),
           py::arg("commit_failures") = 0)
      .def_readwrite("replica_id", &QuorumMember::replica_id)
      .def_readwrite("address", &QuorumMember::address)
      .def_readwrite("store_address", &QuorumMember::store_address)
      .def_readwrite("step", &QuorumMember::step)
      .def_readwrite("world_size", &QuorumMember::world_size)
      .def_readwrite("shrink_only", &QuorumMember::shrink_only)
      .def_readwrite("commit_failures", &QuorumMember::commit_failures)
      .def_property(
          "data", [](const QuorumMember& q) { return json_loads(q.data); },
          [](QuorumMember& q, const py::object& o) { q.data = json_dumps(o); })
      .def("__repr__", [](const QuorumMember& q) {
        return "QuorumMember(replica_id='" + q.replica_id + "', step=" + std::to_string(q.step) + ")";
      });

  py::class_<Quorum>(m, "Quorum")
      .def(py::init<>())
      .def_readwrite("quorum_id", &Quorum::quorum_id)
      .def_readwrite("participants", &Quorum::participants)
      .def_property(
          "created",
          [](const Quorum& q) {
            Timestamp t;
            t.seconds = q.created_ms / 1000;
            t.nanos = (int32_t)((q.created_ms % 1000) * 1000000);
            return t;
          },
          [](Quorum& q, const Timestamp& t) { q.created_ms = t.seconds * 1000 + t.nanos / 1000000; });

  py::class_<QuorumResult>(m, "QuorumResult")
      .def(py::init<>())
      .def_readwrite("quorum_id", &QuorumResult::quorum_id)
      .def_readwrite("replica_rank", &QuorumResult::replica_rank)
      .def_readwrite("replica_world_size", &QuorumResult::replica_world_size)
      .def_readwrite("recover_src_manager_address", &QuorumResult::recover_src_manager_address)
      .def_readwrite("recover_src_replica_rank", &QuorumResult::recover_src_replica_rank)
      .def_readwrite("recover_dst_replica_ranks", &QuorumResult::recover_dst_replica_ranks)
      .def_readwrite("store_address", &QuorumResult::store_address)
      .def_readwrite("max_step", &QuorumResult::max_step)
      .def_readwrite("max_replica_rank", &QuorumResult::max_replica_rank)
      .def_readwrite("max_world_size", &QuorumResult::max_world_size)
      .def_readwrite("heal", &QuorumResult::heal)
      .def_readwrite("commit_failures", &QuorumResult::commit_failures)
      .def_readwrite("replica_ids", &QuorumResult::replica_ids);

  py::class_<Lighthouse>(m, "LighthouseServer")
      .def(py::init([](const std::string& bind, uint64_t min_replicas, std::optional<uint64_t> join_timeout_ms,
                       std::optional<uint64_t> quorum_tick_ms, std::optional<uint64_t> heartbeat_timeout_ms,
                       int64_t quorum_id_base) {
             LighthouseOpt o;
             o.bind = bind;
             o.min_replicas = min_replicas;
             // Python-constructed servers default to a 100 ms join timeout (reference: src/lib.rs:628-630)
             o.join_timeout_ms = join_timeout_ms.value_or(100);
             o.quorum_tick_ms = quorum_tick_ms.value_or(100);
             o.heartbeat_timeout_ms = heartbeat_timeout_ms.value_or(5000);
             o.quorum_id_base = quorum_id_base;
             return std::make_unique<Lighthouse>(o);
           }),
           py::arg("bind"), py::arg("min_replicas"), py::arg("join_timeout_ms") = py::none(),
           py::arg("quorum_tick_ms") = py::none(), py::arg("heartbeat_timeout_ms") = py::none(),
           py::arg("quorum_id_base") = 0,
           py::call_guard<py::gil_scoped_release>())
      .def("address", &Lighthouse::address)
      .def("shutdown", &Lighthouse::shutdown, py::call_guard<py::gil_scoped_release>());

  py::class_<LighthouseClient>(m, "LighthouseClient")
      .def(py::init([](const std::string& addr, std::chrono::duration<double> connect_timeout) {
             return std::make_unique<LighthouseClient>(addr, to_ms(connect_timeout));
           }),
           py::arg("addr"), py::arg("connect_timeout"), py::call_guard<py::gil_scoped_release>())
      .def(
          "quorum",
          [](LighthouseClient& c, const std::string& replica_id, std::chrono::duration<double> timeout,
             const std::string& address, const std::string& store_address, int64_t step, uint64_t world_size,
             bool shrink_only, py::object data) {
            QuorumMember q;
            q.replica_id = replica_id;
            q.address = address;
            q.store_address = store_address;
            q.step = step;
            q.world_size = world_size;
            q.shrink_only = shrink_only;
            q.data = json_dumps(data);
            py::gil_scoped_release rel;
            return c.quorum(q, to_ms(timeout));
          },
          py::arg("replica_id"), py::arg("timeout"), py::arg("address") = "", py::arg("store_address") = "",
          py::arg("step") = 0, py::arg("world_size") = 0, py::arg("shrink_only") = false,
          py::arg("data") = py::none())
      .def(
          "heartbeat",
          [](LighthouseClient& c, const std::string& replica_id, std::chrono::duration<double> timeout) {
            py::gil_scoped_release rel;
            c.heartbeat(replica_id, to_ms(timeout));
          },
          py::arg("replica_id"), py::arg("timeout") = std::chrono::duration<double>(5.0));

  py::class_<ManagerServer>(m, "ManagerServer")
      .def(py::init([](const std::string& replica_id, const std::string& lighthouse_addr, const std::string& hostname,
                       const std::string& bind, const std::string& store_addr, uint64_t world_size,
                       std::chrono::duration<double> heartbeat_interval, std::chrono::duration<double> connect_timeout,
                       int64_t quorum_retries) {
             return std::make_unique<ManagerServer>(replica_id, lighthouse_addr, hostname, bind, store_addr, world_size,
                                                    to_ms(heartbeat_interval), to_ms(connect_timeout), quorum_retries);
           }),
           py::arg("replica_id"), py::arg("lighthouse_addr"), py::arg("hostname"), py::arg("bind"),
           py::arg("store_addr"), py::arg("world_size"), py::arg("heartbeat_interval"), py::arg("connect_timeout"),
           py::arg("quorum_retries"), py::call_guard<py::gil_scoped_release>())
      .def("address", &ManagerServer::address)
      .def("shutdown", &ManagerServer::shutdown, py::call_guard<py::gil_scoped_release>());

  py::class_<ManagerClient>(m, "ManagerClient")
      .def(py::init([](const std::string& addr, std::chrono::duration<double> connect_timeout) {
             return std::make_unique<ManagerClient>(addr, to_ms(connect_timeout));
           }),
           py::arg("addr"), py::arg("connect_timeout"), py::call_guard<py::gil_scoped_release>())
      .def(
          "_quorum",
          [](ManagerClient& c, int64_t group_rank, int64_t step, const std::string& checkpoint_metadata,
             bool shrink_only, std::chrono::duration<double> timeout, int64_t commit_failures, bool init_sync) {
            py::gil_scoped_release rel;
            return c.quorum(group_rank, step, checkpoint_metadata, shrink_only, to_ms(timeout), commit_failures,
                            init_sync);
          },
          py::arg("group_rank"), py::arg("step"), py::arg("checkpoint_metadata"), py::arg("shrink_only"),
          py::arg("timeout"), py::arg("commit_failures"), py::arg("init_sync") = true)
      .def(
          "_checkpoint_metadata",
          [](ManagerClient& c, int64_t rank, std::chrono::duration<double> timeout) {
            py::gil_scoped_release rel;
            return c.checkpoint_metadata(rank, to_ms(timeout));
          },
          py::arg("rank"), py::arg("timeout"))
      .def(
          "should_commit",
          [](ManagerClient& c, int64_t group_rank, int64_t step, bool should_commit,
             std::chrono::duration<double> timeout) {
            py::gil_scoped_release rel;
            return c.should_commit(group_rank, step, should_commit, to_ms(timeout));
          },
          py::arg("group_rank"), py::arg("step"), py::arg("should_commit"), py::arg("timeout"))
      .def(
          "_kill", [](ManagerClient& c, const std::string& msg) {
            py::gil_scoped_release rel;
            c.kill(msg);
          },
          py::arg("msg") = "killed");

  // ---- pure decision procedures, exposed for spec tests ----
  m.def(
      "quorum_compute",
      [](int64_t now_ms, const std::vector<std::pair<int64_t, QuorumMember>>& participants,
         const std::map<std::string, int64_t>& heartbeats, const std::optional<Quorum>& prev_quorum,
         uint64_t min_replicas, uint64_t join_timeout_ms, uint64_t heartbeat_timeout_ms) {
        LighthouseState st;
        for (const auto& [joined, mem] : participants) st.participants[mem.replica_id] = ParticipantDetails{joined, mem};
        st.heartbeats = heartbeats;
        st.prev_quorum = prev_quorum;
        LighthouseOpt o;
        o.min_replicas = min_replicas;
        o.join_timeout_ms = join_timeout_ms;
        o.heartbeat_timeout_ms = heartbeat_timeout_ms;
        QuorumDecision d = quorum_compute(now_ms, st, o);
        return py::make_tuple(d.participants, d.reason);
      },
      py::arg("now_ms"), py::arg("participants"), py::arg("heartbeats"), py::arg("prev_quorum") = py::none(),
      py::arg("min_replicas") = 1, py::arg("join_timeout_ms") = 60000, py::arg("heartbeat_timeout_ms") = 5000);
  m.def("quorum_changed", &quorum_changed);
  m.def("compute_quorum_results", &compute_quorum_results, py::arg("replica_id"), py::arg("group_rank"),
        py::arg("quorum"), py::arg("init_sync") = true);
  m.def("backoff_schedule", [](int n) {
    Backoff b;
    b.max_jitter_ms = 0;
    std::vector<double> v;
    for (int i = 0; i < n; ++i) v.push_back(b.next());
    return v;
  });

  m.def(
      "lighthouse_main",
      [](std::optional<std::vector<std::string>> argv) {
        std::vector<std::string> args;
        if (argv) {
          args = *argv;
        } else {
          py::list sys_argv = py::module_::import("sys").attr("argv");
          for (size_t i = 1; i < sys_argv.size(); ++i) args.push_back(py::cast<std::string>(sys_argv[i]));
        }
        // restore default SIGINT so Ctrl-C stops the server (reference: src/lib.rs:321-347)
        std::signal(SIGINT, SIG_DFL);
        py::gil_scoped_release rel;
        return run_lighthouse_cli(args);
      },
      py::arg("argv") = py::none());
}
