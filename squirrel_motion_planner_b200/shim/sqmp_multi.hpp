// sqmp_multi.hpp — many planning queries per GPU in lock step (BASELINE.json configs[4]; SURVEY.md §7 item 5: "config 5
// needs many concurrent queries per GPU to fill 148 SMs").
//
// A single BiRRT* query asks the library for ~6 tiny answers per iteration (one nearest node, 21..1000 states, ...), each
// a synchronous launch: the GPU idles.  Here B queries of a batch run as cooperative fibers (ucontext) over the SAME
// planner code (sqmp_birrt.hpp, unchanged): a query that needs an answer posts its request and yields; when every live
// query has posted, the scheduler issues ONE library call per request kind for the whole batch —
//   nearest node          -> sqmp_nn_nearest_multi   (every query against its own tree)
//   near vertices         -> sqmp_nn_radius_multi
//   candidate edges       -> sqmp_check_edges        (all queries' candidate edges concatenated)
//   isConfigValid         -> sqmp_check_states
// — hands the answers back and resumes the fibers.  Tree writes need no answer and go straight to the library
// (stream-ordered).  Every query receives exactly the values it would receive alone, so its decisions — and its trace
// hash — are those of the sequential run (tests/test_planner_frontend.py).
#pragma once

#if defined(__x86_64__) && defined(__linux__) && !defined(SQMP_FIBER_UCONTEXT)
#define SQMP_FIBER_ASM 1
#else
#include <ucontext.h>
#endif

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "sqmp_birrt.hpp"
#include "sqmp_shim.hpp"

#ifdef SQMP_FIBER_ASM
// Fiber switch for x86-64 System V: the callee-saved registers, MXCSR and the x87 control word go on the old stack, the
// stack pointers are exchanged, the same state comes off the new stack.  ~10 ns; glibc's swapcontext also saves and
// restores the signal mask with two system calls (~300 ns), which was the largest host cost of a scheduler round.
// (-DSQMP_FIBER_UCONTEXT selects the portable ucontext version.)
extern "C" void sqmp_fiber_switch(void **from_sp, void *to_sp);
extern "C" void sqmp_fiber_boot();
asm(".text\n"
    ".weak sqmp_fiber_switch\n"
    ".type sqmp_fiber_switch,@function\n"
    "sqmp_fiber_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  subq $8, %rsp\n  stmxcsr (%rsp)\n  fnstcw 4(%rsp)\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  ldmxcsr (%rsp)\n  fldcw 4(%rsp)\n  addq $8, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size sqmp_fiber_switch,.-sqmp_fiber_switch\n"
    // first activation of a fiber: rbx = argument, r12 = entry function (placed in the initial frame)
    ".weak sqmp_fiber_boot\n"
    ".type sqmp_fiber_boot,@function\n"
    "sqmp_fiber_boot:\n"
    "  movq %rbx, %rdi\n"
    "  subq $8, %rsp\n"
    "  callq *%r12\n"
    "  ud2\n"
    ".size sqmp_fiber_boot,.-sqmp_fiber_boot\n");
#endif

namespace sqmp_planner {

#ifdef SQMP_FIBER_ASM
// initial frame of a fiber on its own stack: [mxcsr | x87 cw][r15][r14][r13][r12 = entry][rbx = arg][rbp][ret = boot][pad]
inline void *fiber_prepare(char *stack, size_t bytes, void (*entry)(void *), void *arg) {
  uintptr_t top = (reinterpret_cast<uintptr_t>(stack) + bytes) & ~uintptr_t(15);
  uint64_t *sp = reinterpret_cast<uint64_t *>(top);
  *--sp = 0;                                              // pad: after boot's `ret` slot is popped, rsp % 16 == 8
  *--sp = reinterpret_cast<uint64_t>(&sqmp_fiber_boot);   // return address of the first switch
  *--sp = 0;                                              // rbp
  *--sp = reinterpret_cast<uint64_t>(arg);                // rbx
  *--sp = reinterpret_cast<uint64_t>(entry);              // r12
  *--sp = 0;                                              // r13
  *--sp = 0;                                              // r14
  *--sp = 0;                                              // r15
  *--sp = 0x0000037F00001F80ull;                          // MXCSR = 0x1F80 (low 4 bytes), x87 control word = 0x037F
  return sp;
}
#endif

// Where a scheduler's wall time goes (SQMP_PLAN_PROFILE=1 prints it): the batched library calls per kind, the direct
// per-query calls made from inside the fibers (tree writes, sample pools) and — the remainder — the planner's host logic.
struct MultiProfile {
  enum { NEAREST, NEAR, EDGES, VALID, APPEND, UPDATE, SAMPLE, INFORMED, N };
  double sec[N] = {0, 0, 0, 0, 0, 0, 0, 0};
  long calls[N] = {0, 0, 0, 0, 0, 0, 0, 0};
  bool on = false;
  typedef std::chrono::steady_clock clock;
  struct Scope {
    MultiProfile *p; int k; clock::time_point t0;
    Scope(MultiProfile *p_, int k_) : p(p_ && p_->on ? p_ : NULL), k(k_) { if (p) t0 = clock::now(); }
    ~Scope() { if (p) { p->sec[k] += std::chrono::duration<double>(clock::now() - t0).count(); p->calls[k] += 1; } }
  };
  void print(double total, long rounds) const {
    static const char *name[N] = {"nearest_multi", "radius_multi", "check_edges", "check_states", "tree_append", "tree_update", "sample", "sample_informed"};
    double sum = 0.0;
    for (int i = 0; i < N; ++i) sum += sec[i];
    std::fprintf(stderr, "# profile: %.3f s, %ld rounds; host logic + scheduling %.3f s\n", total, rounds, total - sum);
    for (int i = 0; i < N; ++i) std::fprintf(stderr, "#   %-16s %9ld calls %8.3f s %7.1f us/call\n", name[i], calls[i], sec[i], calls[i] ? 1e6 * sec[i] / calls[i] : 0.0);
  }
};

struct MultiRequest {
  enum Kind { NONE = 0, NEAREST, NEAR, EDGES, VALID } kind = NONE;
  // NEAREST / NEAR / VALID
  int slot = 0;
  const double *q = nullptr;
  int exclude = -1;
  double r = 0.0;
  int k = 0;
  // EDGES
  const std::vector<double> *a = nullptr, *b = nullptr;
  int num_points = 0;
  // answers
  int id = 0;
  bool valid = false;
  std::vector<int> *low = nullptr, *high = nullptr, *first_invalid = nullptr;
  int count = 0;
};

// What the scheduler needs from "the library": the batched calls and the tree writes.  GpuService is the product (the
// C ABI of libsqmp_b200.so); the test twin over the CPU oracle lives outside the package (oracle/plan_oracle.cpp).
// Every method returns 0 on success.
class GpuService {
 public:
  GpuService(sqmp_shim::CollisionChecker *checker, bool check_self, bool check_map)
      : ctx_(checker->context()), flags_(sqmp_shim::CollisionChecker::flags(check_self, check_map)) {}
  static int max_batch() { return SQMP_MAX_TREES / 2; }
  unsigned flags() const { return flags_; }
  int tree_reset(int tree) { return sqmp_tree_reset(ctx_, tree); }
  int tree_append(int tree, const double *q, double cost, int32_t *id) { return sqmp_tree_append(ctx_, tree, q, cost, id); }
  int tree_update(int tree, int id, const double *q, double cost) { return sqmp_tree_update(ctx_, tree, id, q, cost); }
  int sample(uint64_t seed, long long first, int n, const double *lo, const double *hi, int max_tries, double *q_out) {
    return sqmp_sample_configs(ctx_, seed, first, n, lo, hi, max_tries, q_out, NULL);
  }
  int sample_informed(uint64_t seed, long long first, int n, const double *lo, const double *hi, int max_tries, const sqmp_informed &el,
                      double *q_out) {
    return sqmp_sample_informed(ctx_, seed, first, n, lo, hi, max_tries, &el, q_out, NULL);
  }
  int nearest_multi(const int32_t *tree, const double *q, int64_t n, int32_t *ids) { return sqmp_nn_nearest_multi(ctx_, tree, q, n, ids); }
  int radius_multi(const int32_t *tree, const double *q, const int32_t *excl, int64_t n, double r, int k, int32_t *low,
                   int32_t *high, int32_t *cnt) {
    return sqmp_nn_radius_multi(ctx_, tree, q, excl, n, r, k, low, high, cnt);
  }
  int check_edges(const double *a, const double *b, int64_t E, int num_points, uint32_t *mask, int32_t *first) {
    return sqmp_check_edges(ctx_, a, b, E, num_points, flags_, mask, first);
  }
  int check_states(const double *q, int64_t n, uint32_t *mask) { return sqmp_check_states(ctx_, q, n, flags_, mask); }
  void report(int rc) const { std::fprintf(stderr, "libsqmp_b200: %s (%s)\n", sqmp_strerror(rc), sqmp_last_error(ctx_)); }

 private:
  sqmp_ctx *ctx_;
  unsigned flags_;
};

template <class Service>
class MultiSchedulerT;

// The Backend of BiRRTStar for ONE query of the batch (fiber `index`: tree slots 2*index and 2*index + 1).
template <class Service>
class FiberBackendT {
 public:
  FiberBackendT() : sched_(NULL), svc_(NULL), prof_(NULL), index_(0), flags_(0), failed_(false) {}
  void bind(MultiSchedulerT<Service> *s, Service *svc, int index, MultiProfile *prof = NULL) {
    sched_ = s; svc_ = svc; index_ = index; flags_ = svc->flags(); prof_ = prof;
  }

  void tree_reset(int slot) { ok(svc_->tree_reset(2 * index_ + slot)); }
  int tree_append(int slot, const double *q, double cost) {
    int32_t id = -1;
    MultiProfile::Scope sc(prof_, MultiProfile::APPEND);
    ok(svc_->tree_append(2 * index_ + slot, q, cost, &id));
    return id;
  }
  void tree_update(int slot, int id, const double *q, double cost) {
    MultiProfile::Scope sc(prof_, MultiProfile::UPDATE);
    ok(svc_->tree_update(2 * index_ + slot, id, q, cost));
  }
  void sample_batch(uint64_t seed, long long first, int n, const double *lo, const double *hi, int max_tries, double *q_out) {
    MultiProfile::Scope sc(prof_, MultiProfile::SAMPLE);
    ok(svc_->sample(seed, first, n, lo, hi, max_tries, q_out));   // once per 256 iterations: not batched
  }
  void sample_informed_batch(uint64_t seed, long long first, int n, const double *lo, const double *hi, int max_tries,
                             const sqmp_informed &el, double *q_out) {
    MultiProfile::Scope sc(prof_, MultiProfile::INFORMED);
    ok(svc_->sample_informed(seed, first, n, lo, hi, max_tries, el, q_out));   // redrawn when the best cost changes: not batched
  }
  inline int nearest(int slot, const double *q);
  inline void near(int slot, const double *q, int exclude_id, double r, int k, std::vector<int> &low, std::vector<int> &high, int &count);
  inline bool config_valid(const double *q);
  inline void check_edges(const std::vector<double> &a, const std::vector<double> &b, int num_points, std::vector<int> &first_invalid);
  bool failed() const { return failed_; }

  MultiRequest req;

 private:
  friend class MultiSchedulerT<Service>;
  void ok(int rc) {
    if (rc != 0) {
      failed_ = true;
      svc_->report(rc);
    }
  }
  MultiSchedulerT<Service> *sched_;
  Service *svc_;
  MultiProfile *prof_;
  int index_;
  unsigned flags_;
  bool failed_;
};

template <class Service>
class MultiSchedulerT {
 public:
  typedef FiberBackendT<Service> FiberBackend;
  struct Result {
    bool valid_endpoints = false;
    Stats stats;
  };

  MultiSchedulerT(Service *service, int batch)
      : svc_(service), flags_(service->flags()),
        batch_(batch < 1 ? 1 : (batch > Service::max_batch() ? Service::max_batch() : batch)), failed_(false), rounds_(0), calls_(0) {
    const char *e = std::getenv("SQMP_PLAN_PROFILE");
    prof_.on = e && e[0] == '1';
  }
  const MultiProfile &profile() const { return prof_; }

  // queries: rows start[8] goal[8]; results in query order.  Returns false on a library error.
  bool run(const double *queries, size_t n_queries, long iterations, uint64_t seed, size_t first_index, const Params &base,
           std::vector<Result> &out) {
    out.assign(n_queries, Result());
    for (size_t g0 = 0; g0 < n_queries && !failed_; g0 += size_t(batch_)) {
      const size_t nb = n_queries - g0 < size_t(batch_) ? n_queries - g0 : size_t(batch_);
      fibers_.assign(nb, Fiber());
      for (size_t i = 0; i < nb; ++i) {
        Fiber &f = fibers_[i];
        f.owner = this;
        f.be.bind(this, svc_, int(i), &prof_);
        f.P = base;
        f.P.seed = seed + first_index + g0 + i;
        f.q = queries + (g0 + i) * 16;
        f.iterations = iterations;
        f.stack.resize(STACK_BYTES);
#ifdef SQMP_FIBER_ASM
        f.sp = fiber_prepare(f.stack.data(), f.stack.size(), &MultiSchedulerT<Service>::entry_ptr, &f);
#else
        getcontext(&f.uc);
        f.uc.uc_stack.ss_sp = f.stack.data();
        f.uc.uc_stack.ss_size = f.stack.size();
        f.uc.uc_link = &main_;
        const uintptr_t p = reinterpret_cast<uintptr_t>(&f);
        makecontext(&f.uc, reinterpret_cast<void (*)()>(&MultiSchedulerT<Service>::entry), 2, unsigned(p & 0xffffffffu), unsigned(p >> 32));
#endif
      }
      size_t alive = nb;
      while (alive > 0 && !failed_) {
        for (size_t i = 0; i < nb; ++i) {   // run every live query up to its next request
          Fiber &f = fibers_[i];
          if (f.done) continue;
          current_ = &f;
#ifdef SQMP_FIBER_ASM
          sqmp_fiber_switch(&main_sp_, f.sp);
#else
          swapcontext(&main_, &f.uc);
#endif
          if (f.done) {
            --alive;
            failed_ = failed_ || f.be.failed();
          }
        }
        serve(nb);
        rounds_ += 1;
      }
      for (size_t i = 0; i < nb; ++i) {
        out[g0 + i].valid_endpoints = fibers_[i].valid_endpoints;
        out[g0 + i].stats = fibers_[i].stats;
      }
    }
    return !failed_;
  }
#ifdef SQMP_FIBER_ASM
  void yield() { sqmp_fiber_switch(&current_->sp, main_sp_); }   // called on a fiber's stack
#else
  void yield() { swapcontext(&current_->uc, &main_); }   // called on a fiber's stack
#endif
  long rounds() const { return rounds_; }
  long library_calls() const { return calls_; }

 private:
  struct Fiber {
    MultiSchedulerT<Service> *owner = NULL;
#ifdef SQMP_FIBER_ASM
    void *sp = NULL;
#else
    ucontext_t uc;
#endif
    std::vector<char> stack;
    FiberBackend be;
    Params P;
    const double *q = NULL;
    long iterations = 0;
    bool done = false, valid_endpoints = false;
    Stats stats;
  };
  static const size_t STACK_BYTES = 256 * 1024;

#ifdef SQMP_FIBER_ASM
  static void entry_ptr(void *p) {
    Fiber *f = static_cast<Fiber *>(p);
    run_query(f);
    sqmp_fiber_switch(&f->sp, f->owner->main_sp_);   // never resumed: the scheduler skips finished fibers
  }
#else
  static void entry(unsigned lo, unsigned hi) {
    run_query(reinterpret_cast<Fiber *>(uintptr_t(lo) | (uintptr_t(hi) << 32)));   // returning resumes uc_link = the scheduler
  }
#endif
  static void run_query(Fiber *f) {
    {
      BiRRTStar<FiberBackend> planner(f->be, f->P);
      f->valid_endpoints = planner.init(f->q, f->q + 8);
      if (f->valid_endpoints) f->stats = planner.run(f->iterations);
    }
    f->be.req.kind = MultiRequest::NONE;
    f->done = true;
  }

  void ok(int rc) {
    if (rc != 0) {
      failed_ = true;
      svc_->report(rc);
    }
  }

  // one library call per request kind present in this round
  void serve(size_t nb) {
    // ---- nearest ----
    idx_.clear(); trees_.clear(); qs_.clear();
    for (size_t i = 0; i < nb; ++i) {
      MultiRequest &r = fibers_[i].be.req;
      if (fibers_[i].done || r.kind != MultiRequest::NEAREST) continue;
      idx_.push_back(int(i));
      trees_.push_back(2 * int(i) + r.slot);
      qs_.insert(qs_.end(), r.q, r.q + 8);
    }
    if (!idx_.empty()) {
      ids_.assign(idx_.size(), 0);
      { MultiProfile::Scope sc(&prof_, MultiProfile::NEAREST);
        ok(svc_->nearest_multi(trees_.data(), qs_.data(), int64_t(idx_.size()), ids_.data())); }
      calls_ += 1;
      for (size_t j = 0; j < idx_.size(); ++j) fibers_[size_t(idx_[j])].be.req.id = ids_[j];
    }
    // ---- near vertices (r and k are planner parameters: equal for the whole batch) ----
    idx_.clear(); trees_.clear(); qs_.clear(); excl_.clear();
    double rr = 0.0;
    int kk = 0;
    for (size_t i = 0; i < nb; ++i) {
      MultiRequest &r = fibers_[i].be.req;
      if (fibers_[i].done || r.kind != MultiRequest::NEAR) continue;
      idx_.push_back(int(i));
      trees_.push_back(2 * int(i) + r.slot);
      qs_.insert(qs_.end(), r.q, r.q + 8);
      excl_.push_back(r.exclude);
      rr = r.r;
      kk = r.k;
    }
    if (!idx_.empty()) {
      low_.assign(idx_.size() * size_t(kk), -1);
      high_.assign(idx_.size() * size_t(kk), -1);
      cnt_.assign(idx_.size(), 0);
      { MultiProfile::Scope sc(&prof_, MultiProfile::NEAR);
        ok(svc_->radius_multi(trees_.data(), qs_.data(), excl_.data(), int64_t(idx_.size()), rr, kk, low_.data(), high_.data(),
                              cnt_.data())); }
      calls_ += 1;
      for (size_t j = 0; j < idx_.size(); ++j) {
        MultiRequest &r = fibers_[size_t(idx_[j])].be.req;
        r.low->assign(low_.begin() + long(j) * kk, low_.begin() + long(j + 1) * kk);
        r.high->assign(high_.begin() + long(j) * kk, high_.begin() + long(j + 1) * kk);
        r.count = cnt_[j];
      }
    }
    // ---- candidate edges ----
    idx_.clear(); ea_.clear(); eb_.clear();
    int np = 0;
    for (size_t i = 0; i < nb; ++i) {
      MultiRequest &r = fibers_[i].be.req;
      if (fibers_[i].done || r.kind != MultiRequest::EDGES) continue;
      idx_.push_back(int(i));
      ea_.insert(ea_.end(), r.a->begin(), r.a->end());
      eb_.insert(eb_.end(), r.b->begin(), r.b->end());
      np = r.num_points;
    }
    if (!idx_.empty()) {
      const int64_t E = int64_t(ea_.size() / 8);
      first_.assign(size_t(E), -1);
      if (flags_ != 0) {
        mask_.assign(size_t((E * (np + 1) + 31) / 32), 0u);
        { MultiProfile::Scope sc(&prof_, MultiProfile::EDGES);
          ok(svc_->check_edges(ea_.data(), eb_.data(), E, np, mask_.data(), first_.data())); }
        calls_ += 1;
      }
      size_t off = 0;
      for (size_t j = 0; j < idx_.size(); ++j) {
        MultiRequest &r = fibers_[size_t(idx_[j])].be.req;
        const size_t e = r.a->size() / 8;
        r.first_invalid->assign(first_.begin() + long(off), first_.begin() + long(off + e));
        off += e;
      }
    }
    // ---- isConfigValid ----
    idx_.clear(); qs_.clear();
    for (size_t i = 0; i < nb; ++i) {
      MultiRequest &r = fibers_[i].be.req;
      if (fibers_[i].done || r.kind != MultiRequest::VALID) continue;
      idx_.push_back(int(i));
      qs_.insert(qs_.end(), r.q, r.q + 8);
    }
    if (!idx_.empty()) {
      mask_.assign((idx_.size() + 31) / 32, 0u);
      if (flags_ != 0) {
        { MultiProfile::Scope sc(&prof_, MultiProfile::VALID);
          ok(svc_->check_states(qs_.data(), int64_t(idx_.size()), mask_.data())); }
        calls_ += 1;
      }
      for (size_t j = 0; j < idx_.size(); ++j)
        fibers_[size_t(idx_[j])].be.req.valid = ((mask_[j >> 5] >> (j & 31)) & 1u) == 0;
    }
    for (size_t i = 0; i < nb; ++i) fibers_[i].be.req.kind = MultiRequest::NONE;
  }

  Service *svc_;
  unsigned flags_;
  int batch_;
  bool failed_;
  long rounds_, calls_;
  MultiProfile prof_;
#ifdef SQMP_FIBER_ASM
  void *main_sp_ = NULL;
#else
  ucontext_t main_;
#endif
  Fiber *current_ = NULL;
  std::vector<Fiber> fibers_;
  std::vector<int> idx_;
  std::vector<int32_t> trees_, excl_, ids_, low_, high_, cnt_, first_;
  std::vector<double> qs_, ea_, eb_;
  std::vector<uint32_t> mask_;
};

template <class Service>
inline int FiberBackendT<Service>::nearest(int slot, const double *q) {
  req.kind = MultiRequest::NEAREST;
  req.slot = slot;
  req.q = q;
  sched_->yield();
  return req.id;
}
template <class Service>
inline void FiberBackendT<Service>::near(int slot, const double *q, int exclude_id, double r, int k, std::vector<int> &low,
                               std::vector<int> &high, int &count) {
  req.kind = MultiRequest::NEAR;
  req.slot = slot;
  req.q = q;
  req.exclude = exclude_id;
  req.r = r;
  req.k = k;
  req.low = &low;
  req.high = &high;
  sched_->yield();
  count = req.count;
}
template <class Service>
inline bool FiberBackendT<Service>::config_valid(const double *q) {
  if (flags_ == 0) return true;
  req.kind = MultiRequest::VALID;
  req.q = q;
  sched_->yield();
  return req.valid;
}
template <class Service>
inline void FiberBackendT<Service>::check_edges(const std::vector<double> &a, const std::vector<double> &b, int num_points,
                                      std::vector<int> &first_invalid) {
  req.kind = MultiRequest::EDGES;
  req.a = &a;
  req.b = &b;
  req.num_points = num_points;
  req.first_invalid = &first_invalid;
  sched_->yield();
}

typedef MultiSchedulerT<GpuService> MultiScheduler;   // the product instantiation

}  // namespace sqmp_planner
