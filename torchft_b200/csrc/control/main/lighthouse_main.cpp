// Standalone lighthouse executable (reference: src/bin/lighthouse.rs).
#include <string>
#include <vector>

int run_lighthouse_cli(const std::vector<std::string>& args);

int main(int argc, char** argv) {
  std::vector<std::string> args(argv + 1, argv + argc);
  return run_lighthouse_cli(args);
}
