// Shared CLI driver for the standalone `torchft_b200_lighthouse` binary and
// `torchft_b200._C.lighthouse_main` (reference flags: src/lighthouse.rs:94-131).
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "lighthouse.h"

namespace {
void usage() {
  fprintf(stderr,
          "torchft_b200_lighthouse --min_replicas N [--bind [::]:29510] [--join_timeout_ms 60000]\n"
          "                        [--quorum_tick_ms 100] [--heartbeat_timeout_ms 5000]\n"
          "                        [--quorum_id_base 0|auto|N]   (auto: clock-based, ids never repeat across restarts)\n");
}
}  // namespace

int run_lighthouse_cli(const std::vector<std::string>& args) {
  tft::LighthouseOpt opt;
  bool have_min = false;
  for (size_t i = 0; i < args.size(); ++i) {
    std::string k = args[i], v;
    size_t eq = k.find('=');
    if (eq != std::string::npos) {
      v = k.substr(eq + 1);
      k = k.substr(0, eq);
    } else if (k == "-h" || k == "--help") {
      usage();
      return 0;
    } else if (i + 1 < args.size()) {
      v = args[++i];
    } else {
      fprintf(stderr, "missing value for %s\n", k.c_str());
      usage();
      return 2;
    }
    try {
      if (k == "--bind") opt.bind = v;
      else if (k == "--min_replicas") { opt.min_replicas = std::stoull(v); have_min = true; }
      else if (k == "--join_timeout_ms") opt.join_timeout_ms = std::stoull(v);
      else if (k == "--quorum_tick_ms") opt.quorum_tick_ms = std::stoull(v);
      else if (k == "--heartbeat_timeout_ms") opt.heartbeat_timeout_ms = std::stoull(v);
      else if (k == "--quorum_id_base") opt.quorum_id_base = (v == "auto") ? -1 : std::stoll(v);
      else {
        fprintf(stderr, "unknown flag %s\n", k.c_str());
        usage();
        return 2;
      }
    } catch (const std::exception&) {
      fprintf(stderr, "bad value for %s: %s\n", k.c_str(), v.c_str());
      return 2;
    }
  }
  if (!have_min) {
    fprintf(stderr, "--min_replicas is required\n");
    usage();
    return 2;
  }
  setenv("TORCHFT_B200_LOG", "info", 0);
  tft::Lighthouse lh(opt);
  fprintf(stderr, "Lighthouse listening on: %s\n", lh.address().c_str());
  while (true) std::this_thread::sleep_for(std::chrono::seconds(3600));
  return 0;
}
