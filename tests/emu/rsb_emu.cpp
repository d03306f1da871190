/*
 * rsb_emu.cpp -- TEST INFRASTRUCTURE: runs the device code of robosuite_benchmark_b200/csrc/rsb_dev.h on the host.
 *
 * The lanes of one group are ucontext fibers; gsync()/shuffles are barrier points where the scheduler switches lanes.
 * Lane order within a barrier interval is configurable (forward / reverse / strided): a missing gsync() shows up as a
 * result that depends on the order.  This exists because the build container has no GPU; it is never loaded by the
 * product path (robosuite_benchmark_b200.backend loads only the CUDA library and fails loudly without it).
 */
#define RSB_EMU
#include <ucontext.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <functional>
#include <algorithm>

#include "../../robosuite_benchmark_b200/csrc/rsb_dev.h"

DevModel emu_model;
float *emu_smem = nullptr;

namespace {
constexpr int NL = RSB_LANES;
constexpr size_t STACK = 1 << 18;

struct Sched {
  ucontext_t main_ctx, lane_ctx[NL];
  std::vector<unsigned char> stacks;
  bool done[NL];
  int cur = -1, order = 0;
  float xf[NL]; int xi[NL];
  std::function<void(int)> body;
} S;

void lane_entry(int lane) { S.body(lane); S.done[lane] = true; swapcontext(&S.lane_ctx[lane], &S.main_ctx); }

void run_group(std::function<void(int)> body) {
  S.body = body;
  if (S.stacks.empty()) S.stacks.resize(STACK * NL);
  for (int l = 0; l < NL; l++) {
    S.done[l] = false; getcontext(&S.lane_ctx[l]);
    S.lane_ctx[l].uc_stack.ss_sp = S.stacks.data() + STACK * (size_t)l; S.lane_ctx[l].uc_stack.ss_size = STACK; S.lane_ctx[l].uc_link = &S.main_ctx;
    makecontext(&S.lane_ctx[l], (void (*)())lane_entry, 1, l);
  }
  for (;;) {
    int ndone = 0;
    for (int k = 0; k < NL; k++) {
      int l = S.order == 0 ? k : (S.order == 1 ? NL - 1 - k : (k * 7 + 3) % NL);
      if (S.done[l]) { ndone++; continue; }
      S.cur = l; swapcontext(&S.main_ctx, &S.lane_ctx[l]);
      if (S.done[l]) ndone++;
    }
    if (ndone == NL) break;
    if (ndone != 0) { fprintf(stderr, "rsb_emu: divergent barrier (some lanes exited while others wait)\n"); abort(); }
  }
}
}  // namespace

void emu_sync() { int l = S.cur; swapcontext(&S.lane_ctx[l], &S.main_ctx); }
float emu_shfl_f(float v, int src) { int l = S.cur; S.xf[l] = v; emu_sync(); float r = S.xf[src & (NL - 1)]; emu_sync(); return r; }
int emu_shfl_i(int v, int src) { int l = S.cur; S.xi[l] = v; emu_sync(); int r = S.xi[src & (NL - 1)]; emu_sync(); return r; }

/* the shared-memory image sits between two guard zones: a write outside the env's slice (which on the GPU would land in the neighbour
   env's slice) trips check_guards() */
constexpr int GUARD = 64;
constexpr uint32_t GUARD_WORD = 0x7fc0dead;
struct EmuEnv {
  RsbHostModel hm; DevModel dm;
  std::vector<float> smem, state, dbg;
  float *slice() { return smem.data() + GUARD; }
  void arm() { for (int i = 0; i < GUARD; i++) { memcpy(&smem[(size_t)i], &GUARD_WORD, 4); memcpy(&smem[smem.size() - 1 - (size_t)i], &GUARD_WORD, 4); } }
  void check_guards() const {
    for (int i = 0; i < GUARD; i++) { uint32_t a, b; memcpy(&a, &smem[(size_t)i], 4); memcpy(&b, &smem[smem.size() - 1 - (size_t)i], 4);
      if (a != GUARD_WORD || b != GUARD_WORD) { fprintf(stderr, "rsb_emu: write outside the env's shared-memory slice (guard word %d)\n", i); abort(); } }
  }
};

static unsigned int g_counters[8];

extern "C" {

void *emu_create(const rsb_model *m, const rsb_task *t, int ncon_max, int nefc_max) {
  EmuEnv *e = new EmuEnv();
  if (!rsb_build_host_model(m, t, ncon_max, nefc_max, e->hm)) { fprintf(stderr, "rsb_emu: %s\n", e->hm.error.c_str()); delete e; return nullptr; }
  e->dm = e->hm.dm; rsb_fixup_pointers(e->dm, e->hm.arena.data());
  memset(g_counters, 0, sizeof g_counters); e->dm.counters = g_counters;
  e->smem.assign((size_t)e->dm.smem_words + 2 * GUARD, 0.0f); e->arm(); e->state.assign((size_t)e->dm.st_words, 0.0f);
  e->dbg.assign((size_t)RSB_DBG_WORDS(e->dm.nv, ncon_max, nefc_max), 0.0f);
  return e;
}
void emu_destroy(void *h) { delete (EmuEnv *)h; }
void emu_set_order(int order) { S.order = order; }
static float g_fill = 0.0f; static int g_do_fill = 0;
/* fill the shared-memory image with a garbage value before every call: a read of a word the step has not written shows up */
void emu_set_fill(float v, int on) { g_fill = v; g_do_fill = on; }
int emu_smem_words(void *h) { return ((EmuEnv *)h)->dm.smem_words; }
int emu_state_words(void *h) { return ((EmuEnv *)h)->dm.st_words; }
int emu_dbg_words(void *h) { return (int)((EmuEnv *)h)->dbg.size(); }
void emu_set_ls_tol(void *h, float tol) { ((EmuEnv *)h)->dm.ls_tol = tol; }
/* event counters of the device code (contact / constraint-row truncation, steps after done), cleared at emu_create */
void emu_counters(void *h, unsigned int *out3) { (void)h; for (int k = 0; k < 3; k++) out3[k] = g_counters[k]; }
void emu_get_solver(void *h, double *out4) { EmuEnv *e = (EmuEnv *)h; out4[0] = e->dm.solver_iters; out4[1] = e->dm.solver_tol; out4[2] = e->dm.ls_iters; out4[3] = e->dm.ls_tol; }
void emu_set_solver(void *h, int iters, int ls_iters, float tol) { EmuEnv *e = (EmuEnv *)h; e->dm.solver_iters = iters; e->dm.ls_iters = ls_iters; e->dm.solver_tol = tol; }

void emu_get_state(void *h, float *out) { EmuEnv *e = (EmuEnv *)h; memcpy(out, e->state.data(), e->state.size() * 4); }
void emu_set_state(void *h, const float *in) { EmuEnv *e = (EmuEnv *)h; memcpy(e->state.data(), in, e->state.size() * 4); }

void emu_reset(void *h, uint64_t seed, uint64_t env_id, float *obs) {
  EmuEnv *e = (EmuEnv *)h;
  emu_model = e->dm; emu_smem = e->slice();
  run_group([&](int lane) { Grp g{lane, (RSB_LANES == 32) ? 0xffffffffu : ((1u << RSB_LANES) - 1u)}; env_reset(0, g, e->state.data(), seed, env_id, obs, true); });
  e->check_guards();
}
int emu_step(void *h, const float *action, float *obs, float *reward) {
  EmuEnv *e = (EmuEnv *)h; unsigned char done = 0;
  if (g_do_fill) { std::fill(e->smem.begin(), e->smem.end(), g_fill); e->arm(); }
  emu_model = e->dm; emu_smem = e->slice();
  run_group([&](int lane) { Grp g{lane, (RSB_LANES == 32) ? 0xffffffffu : ((1u << RSB_LANES) - 1u)}; env_step(0, g, e->state.data(), action, obs, nullptr, reward, &done, nullptr, true); });
  e->check_guards();
  return done;
}
/* one physics substep from the stored state, state written back, internals dumped */
void emu_debug_substep(void *h, const float *action, int policy_step, float *dbg) {
  EmuEnv *e = (EmuEnv *)h;
  emu_model = e->dm; emu_smem = e->slice();
  run_group([&](int lane) {
    Grp g{lane, (RSB_LANES == 32) ? 0xffffffffu : ((1u << RSB_LANES) - 1u)}; const DevModel &m = e->dm; float *s = e->slice();
    load_state(0, e->state.data(), g);
    for (int i = g.lane; i < m.act_dim; i += RSB_LANES) s[m.o_act + i] = action[i];
    gsync(g);
    long long pt_ = 0; substep(0, g, policy_step != 0, pt_);
    dump_debug(0, g, dbg);
    store_state(0, e->state.data(), g);
  });
  e->check_guards();
}
void emu_random_action(void *h, uint64_t seed, uint64_t env_id, uint64_t step, float *action) {
  EmuEnv *e = (EmuEnv *)h; for (int blk = 0; 4 * blk < e->dm.act_dim; blk++) random_action_block(seed, env_id, step, blk, e->dm.act_dim, action);
}
}
