// Paged-KV block manager + continuous-batching scheduler (pure C++, no CUDA: unit-testable on CPU).
//
// * BlockManager: free-list allocator over `num_blocks` KV pages with reference counts, so the N samples
//   of one prompt share its full prompt pages (prefix sharing, the vLLM feature GRPO/RLOO/RAFT rely on).
// * Scheduler: request groups (prompt, n samples, max_new_tokens) wait in FIFO order; `admit()` moves
//   groups into the running set while pages and sequence slots are available.  Two policies:
//     reserve    - every admitted sequence owns all pages it can ever need (prompt + max_new_tokens),
//                  so decode never allocates and the whole decode loop can live in one CUDA graph;
//     on_demand  - pages are allocated at page boundaries during decode; when the pool runs dry the
//                  youngest running group is preempted (pages freed, group re-queued for recompute).
// NRL_RUNTIME_NO_PYBIND: the classes alone (csrc/runtime_asan_test.cpp builds them with -fsanitize=address,undefined)
#ifndef NRL_RUNTIME_NO_PYBIND
#include "runtime.h"

#include <pybind11/numpy.h>
#include <pybind11/stl.h>
#endif

#include <algorithm>
#include <deque>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#ifndef NRL_RUNTIME_NO_PYBIND
namespace py = pybind11;
#endif

namespace nrl {

class BlockManager {
 public:
  BlockManager(int num_blocks, int block_size) : block_size_(block_size), ref_(num_blocks, 0) {
    if (num_blocks <= 0 || block_size <= 0) throw std::invalid_argument("num_blocks and block_size must be positive");
    free_.reserve(num_blocks);
    for (int i = num_blocks - 1; i >= 0; --i) free_.push_back(i);
  }
  int num_free() const { return static_cast<int>(free_.size()); }
  int num_blocks() const { return static_cast<int>(ref_.size()); }
  int block_size() const { return block_size_; }
  int blocks_for(int tokens) const { return (tokens + block_size_ - 1) / block_size_; }

  std::vector<int> allocate(int n) {
    if (n > num_free()) throw std::runtime_error("BlockManager: out of KV pages");
    std::vector<int> out(n);
    for (int i = 0; i < n; ++i) {
      out[i] = free_.back();
      free_.pop_back();
      ref_[out[i]] = 1;
    }
    return out;
  }
  void incref(const std::vector<int>& blocks) {
    for (int b : blocks) {
      if (ref_.at(b) <= 0) throw std::runtime_error("incref on a free page");
      ++ref_[b];
    }
  }
  void release(const std::vector<int>& blocks) {
    for (int b : blocks) {
      if (ref_.at(b) <= 0) throw std::runtime_error("double free of a KV page");
      if (--ref_[b] == 0) free_.push_back(b);
    }
  }
  int refcount(int b) const { return ref_.at(b); }

 private:
  int block_size_;
  std::vector<int> ref_;
  std::vector<int> free_;
};

struct Seq {
  int id = -1, group = -1, sample = 0;
  int prompt_len = 0, num_generated = 0, max_new = 0;
  bool finished = false;
  std::vector<int> blocks;     // block table (shared prompt pages first)
  int num_tokens() const { return prompt_len + num_generated; }
};

struct Group {
  int id = -1, prompt_len = 0, n = 1, max_new = 0;
  std::vector<int> seq_ids;
  bool running = false;
  int admissions = 0;          // > 1 means it was preempted and recomputed
};

class Scheduler {
 public:
  Scheduler(int num_blocks, int block_size, int max_num_seqs, bool reserve)
      : bm_(num_blocks, block_size), max_num_seqs_(max_num_seqs), reserve_(reserve) {}

  int add_request(int prompt_len, int n, int max_new_tokens) {
    if (prompt_len <= 0 || n <= 0 || max_new_tokens <= 0) throw std::invalid_argument("bad request");
    Group g;
    g.id = static_cast<int>(groups_.size());
    g.prompt_len = prompt_len; g.n = n; g.max_new = max_new_tokens;
    for (int i = 0; i < n; ++i) {
      Seq s;
      s.id = static_cast<int>(seqs_.size());
      s.group = g.id; s.sample = i; s.prompt_len = prompt_len; s.max_new = max_new_tokens;
      g.seq_ids.push_back(s.id);
      seqs_.push_back(s);
    }
    int need = pages_needed(g);
    if (need > bm_.num_blocks()) throw std::runtime_error("request can never fit in the KV pool");
    groups_.push_back(g);
    waiting_.push_back(g.id);
    return g.id;
  }

  // pages a group needs at admission
  int pages_needed(const Group& g) const {
    const int bs = bm_.block_size();
    const int shared = g.prompt_len / bs;                       // full prompt pages, shared by the n samples
    const int total_tokens = reserve_ ? g.prompt_len + g.max_new : g.prompt_len + 1;
    const int per_seq_total = bm_.blocks_for(total_tokens);
    return shared + g.n * (per_seq_total - shared);
  }

  // Admit waiting groups (FIFO) while resources last.  Returns the admitted group ids.
  std::vector<int> admit() {
    std::vector<int> out;
    while (!waiting_.empty()) {
      Group& g = groups_[waiting_.front()];
      if (num_running_seqs_ + g.n > max_num_seqs_ && num_running_seqs_ > 0) break;
      if (pages_needed(g) > bm_.num_free()) break;
      waiting_.pop_front();
      const int bs = bm_.block_size();
      const int shared = g.prompt_len / bs;
      std::vector<int> shared_pages = bm_.allocate(shared);
      const int total_tokens = reserve_ ? g.prompt_len + g.max_new : g.prompt_len + 1;
      const int priv = bm_.blocks_for(total_tokens) - shared;
      for (int i = 0; i < g.n; ++i) {
        Seq& s = seqs_[g.seq_ids[i]];
        s.num_generated = 0;
        s.finished = false;
        s.blocks = shared_pages;
        if (i > 0) bm_.incref(shared_pages);
        std::vector<int> p = bm_.allocate(priv);
        s.blocks.insert(s.blocks.end(), p.begin(), p.end());
      }
      g.running = true;
      ++g.admissions;
      num_running_seqs_ += g.n;
      running_.push_back(g.id);
      out.push_back(g.id);
    }
    return out;
  }

  // Account for one decoded token per running, unfinished sequence.  In on_demand mode this allocates
  // pages at page boundaries and may preempt; returns the preempted group ids.
  std::vector<int> advance(const std::vector<int>& seq_ids) {
    std::vector<int> preempted;
    for (int sid : seq_ids) {
      Seq& s = seqs_.at(sid);
      if (s.finished || !groups_[s.group].running) continue;
      ++s.num_generated;
      if (reserve_) continue;
      const int need = bm_.blocks_for(s.num_tokens() + 1);       // room for the next token's KV
      while (static_cast<int>(s.blocks.size()) < need) {
        if (bm_.num_free() == 0) {
          int victim = pick_victim(s.group);
          if (victim < 0) throw std::runtime_error("KV pool exhausted with nothing left to preempt");
          preempt(victim);
          preempted.push_back(victim);
          if (victim == s.group) break;
          continue;
        }
        std::vector<int> p = bm_.allocate(1);
        s.blocks.push_back(p[0]);
      }
    }
    return preempted;
  }

  void finish(const std::vector<int>& seq_ids) {
    for (int sid : seq_ids) {
      Seq& s = seqs_.at(sid);
      if (s.finished) continue;
      s.finished = true;
      Group& g = groups_[s.group];
      if (!g.running) continue;
      bm_.release(s.blocks);
      s.blocks.clear();
      --num_running_seqs_;
      bool all = true;
      for (int q : g.seq_ids) all = all && seqs_[q].finished;
      if (all) {
        g.running = false;
        running_.erase(std::remove(running_.begin(), running_.end(), g.id), running_.end());
      }
    }
  }

  // ---- queries ----
  std::vector<int> group_seqs(int gid) const { return groups_.at(gid).seq_ids; }
  std::vector<int> block_table(int sid) const { return seqs_.at(sid).blocks; }
  // Row-major [sids.size(), width] page table of a whole batch (unused entries = fill) written to `out`: one call per
  // re-batch instead of a Python loop over sequences.
  void block_tables_into(const std::vector<int>& sids, int width, int fill, int* out) const {
    for (size_t r = 0; r < sids.size(); ++r) {
      const std::vector<int>& b = seqs_.at(sids[r]).blocks;
      if (static_cast<int>(b.size()) > width) throw std::runtime_error("block_tables: a sequence holds more pages than the table is wide");
      int* row = out + r * static_cast<size_t>(width);
      std::copy(b.begin(), b.end(), row);
      std::fill(row + b.size(), row + width, fill);
    }
  }
  int seq_len(int sid) const { return seqs_.at(sid).num_tokens(); }
  int num_shared_pages(int gid) const { return groups_.at(gid).prompt_len / bm_.block_size(); }
  int num_waiting() const { return static_cast<int>(waiting_.size()); }
  int num_running_seqs() const { return num_running_seqs_; }
  int num_free_blocks() const { return bm_.num_free(); }
  int admissions(int gid) const { return groups_.at(gid).admissions; }
  bool is_finished(int sid) const { return seqs_.at(sid).finished; }
  std::vector<int> running_groups() const { return running_; }
  BlockManager& block_manager() { return bm_; }

  // Prefill write plan of a group: (token index within the prompt, slot) pairs.  Tokens on shared full
  // pages are written once; the trailing partial page is written once per sample (private copies).
  std::pair<std::vector<int>, std::vector<int>> prefill_slots(int gid) const {
    const Group& g = groups_.at(gid);
    const int bs = bm_.block_size();
    const int shared_tokens = (g.prompt_len / bs) * bs;
    std::vector<int> tok, slot;
    const Seq& s0 = seqs_[g.seq_ids[0]];
    for (int t = 0; t < shared_tokens; ++t) {
      tok.push_back(t);
      slot.push_back(s0.blocks[t / bs] * bs + t % bs);
    }
    for (int sid : g.seq_ids) {
      const Seq& s = seqs_[sid];
      for (int t = shared_tokens; t < g.prompt_len; ++t) {
        tok.push_back(t);
        slot.push_back(s.blocks[t / bs] * bs + t % bs);
      }
    }
    return {tok, slot};
  }

 private:
  int pick_victim(int requester) const {
    // youngest running group that is not the requester; the requester itself as a last resort
    for (auto it = running_.rbegin(); it != running_.rend(); ++it)
      if (*it != requester) return *it;
    return running_.empty() ? -1 : running_.back();
  }
  void preempt(int gid) {
    Group& g = groups_[gid];
    for (int sid : g.seq_ids) {
      Seq& s = seqs_[sid];
      if (!s.finished) {
        bm_.release(s.blocks);
        --num_running_seqs_;
      }
      s.blocks.clear();
      s.num_generated = 0;
      s.finished = false;
    }
    g.running = false;
    running_.erase(std::remove(running_.begin(), running_.end(), gid), running_.end());
    waiting_.push_front(gid);
  }

  BlockManager bm_;
  int max_num_seqs_;
  bool reserve_;
  int num_running_seqs_ = 0;
  std::vector<Seq> seqs_;
  std::vector<Group> groups_;
  std::deque<int> waiting_;
  std::vector<int> running_;
};

#ifndef NRL_RUNTIME_NO_PYBIND
void bind_runtime(py::module_& m) {
  py::class_<BlockManager>(m, "BlockManager")
      .def(py::init<int, int>(), py::arg("num_blocks"), py::arg("block_size"))
      .def("num_free", &BlockManager::num_free)
      .def("num_blocks", &BlockManager::num_blocks)
      .def("block_size", &BlockManager::block_size)
      .def("blocks_for", &BlockManager::blocks_for)
      .def("allocate", &BlockManager::allocate)
      .def("incref", &BlockManager::incref)
      .def("release", &BlockManager::release)
      .def("refcount", &BlockManager::refcount);
  py::class_<Scheduler>(m, "Scheduler")
      .def(py::init<int, int, int, bool>(), py::arg("num_blocks"), py::arg("block_size"), py::arg("max_num_seqs"),
           py::arg("reserve") = true)
      .def("add_request", &Scheduler::add_request)
      .def("admit", &Scheduler::admit)
      .def("advance", &Scheduler::advance)
      .def("finish", &Scheduler::finish)
      .def("group_seqs", &Scheduler::group_seqs)
      .def("block_table", &Scheduler::block_table)
      .def("block_tables", [](const Scheduler& self, const std::vector<int>& sids, int width, int fill) {
             py::array_t<int> out({static_cast<py::ssize_t>(sids.size()), static_cast<py::ssize_t>(width)});
             self.block_tables_into(sids, width, fill, out.mutable_data());
             return out;
           }, py::arg("seq_ids"), py::arg("width"), py::arg("fill"))
      .def("seq_len", &Scheduler::seq_len)
      .def("num_shared_pages", &Scheduler::num_shared_pages)
      .def("num_waiting", &Scheduler::num_waiting)
      .def("num_running_seqs", &Scheduler::num_running_seqs)
      .def("num_free_blocks", &Scheduler::num_free_blocks)
      .def("admissions", &Scheduler::admissions)
      .def("is_finished", &Scheduler::is_finished)
      .def("running_groups", &Scheduler::running_groups)
      .def("prefill_slots", &Scheduler::prefill_slots);
}
#endif

}  // namespace nrl
