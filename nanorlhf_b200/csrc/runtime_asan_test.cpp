// Host-sanitizer target: the sampler's C++ runtime (paged-KV BlockManager + continuous-batching
// Scheduler) driven through a randomized admit / advance / finish / preempt workload under
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -DNRL_RUNTIME_NO_PYBIND runtime_asan_test.cpp
// Invariants checked every step: no page is owned twice, free + owned == total, reference counts match the tables.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <set>

#ifndef NRL_RUNTIME_NO_PYBIND
#define NRL_RUNTIME_NO_PYBIND
#endif
#include "runtime.cpp"

using namespace nrl;

static void check(Scheduler& s, int total_blocks) {
  std::map<int, int> uses;
  for (int g : s.running_groups())
    for (int sid : s.group_seqs(g))
      if (!s.is_finished(sid))
        for (int b : s.block_table(sid)) {
          if (b < 0 || b >= total_blocks) { std::fprintf(stderr, "block id %d out of range\n", b); std::abort(); }
          ++uses[b];
        }
  if (static_cast<int>(uses.size()) + s.num_free_blocks() != total_blocks) {
    std::fprintf(stderr, "page accounting broken: %zu owned + %d free != %d\n", uses.size(), s.num_free_blocks(), total_blocks);
    std::abort();
  }
}

int main() {
  std::mt19937 rng(1234);
  for (int reserve = 0; reserve < 2; ++reserve) {
    const int total = 257, page = 16;
    Scheduler s(total, page, 64, reserve != 0);
    std::vector<int> groups;
    for (int i = 0; i < 40; ++i) groups.push_back(s.add_request(1 + rng() % 90, 1 + rng() % 4, 8 + rng() % 120));
    std::set<int> live;
    int steps = 0;
    while (steps++ < 20000) {
      for (int g : s.admit())
        for (int sid : s.group_seqs(g)) live.insert(sid);
      check(s, total);
      if (live.empty() && s.num_waiting() == 0) break;
      std::vector<int> ids(live.begin(), live.end());
      for (int g : s.advance(ids))                       // on_demand mode may preempt: those sequences restart later
        for (int sid : s.group_seqs(g)) live.erase(sid);
      std::vector<int> done;
      for (int sid : ids)
        if (live.count(sid) && (rng() % 37 == 0 || s.seq_len(sid) > 200)) done.push_back(sid);
      s.finish(done);
      s.finish(done);                                    // idempotent
      for (int sid : done) live.erase(sid);
      check(s, total);
    }
    if (s.num_free_blocks() != total || s.num_running_seqs() != 0) {
      std::fprintf(stderr, "leak: %d free of %d, %d running\n", s.num_free_blocks(), total, s.num_running_seqs());
      return 1;
    }
    std::printf("mode %s: %d scheduler steps, all %d pages returned\n", reserve ? "reserve" : "on_demand", steps, total);
  }
  return 0;
}
