// peer_protocol_model.cc -- a CPU model of the synchronisation skeleton of k_peer_loop
// (or-tools_b200/csrc/device_ops.cu), built with -fsanitize=thread by
// tests/test_peer_protocol_model.py. TEST INFRASTRUCTURE: it models the protocol, it is not the
// CUDA code. G "ranks" of B "blocks" (one thread each) run A attempts of
//
//   P   slice step: x' and x~ of the rank's column slice, x~ stored into EVERY arena     | grid + peer barrier A
//   D   row slices handed out by tickets: K[R_g,:] x~ read from the own arena, dual update,
//       MODE 0: y' stored into every arena; per-slice sums in per-slice slots            | MODE 1: plain grid barrier,
//   T1  (MODE 1) column slices by tickets: (K[R_g,:])^T y' into the own arena's partial  |   block sums over fixed ranges
//       last block: sums in a fixed order, the triple stored into every arena            | grid + peer barrier B
//   T   one block takes the decision from the G triples and writes the OTHER state slot;
//       MODE 0: (K[:,C_g])^T y' from the all-gathered y'; MODE 1: pull of every rank's partial | plain grid barrier
//
// with exactly the barriers of the kernel: a grid barrier is a ticket (fetch_add) plus a generation
// word; the block that draws the last ticket does the cross-rank part (epoch flags stored into the
// peers' arenas, wait for the peers' flags in the own arena) before it releases the grid. All
// vectors are PLAIN doubles: ThreadSanitizer then checks what DESIGN.md 5 argues -- that two
// cross-rank barriers per attempt order every write into an arena after the last read of the
// previous contents, for tickets handed out in any order -- and the result is compared bitwise with
// a sequential run of the same arithmetic. The device's fences become release / acquire operations
// on the same words (TSan does not model stand-alone fences); -DDROP_PEER_A / -DDROP_PEER_B remove
// the cross-rank part of one barrier and must make the run fail (negative controls). The rounds of the
// trust-region solve follow the attempts (tr_main; -DTR_SINGLE_SLOT is their negative control).
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#ifndef MODE
#define MODE 0
#endif

namespace {
constexpr int G = 3, B = 3, A = 40;      // ranks, blocks per rank, attempts
constexpr int N = 24, M = 18;            // columns, rows (divisible by G)
constexpr int NS = N / G, MS = M / G;    // slice / row block per rank
constexpr int kRowSlice = 2;             // rows per ticket of D
constexpr int kColSlice = 4;             // columns per ticket of T1 / T

struct State { int cur, cand, halt; double step; };

struct Arena {  // what the peers write into (PeerLayout)
  double xt[N];
  double y[M];
  double partial[N];
  double scal[4 * G];
  double tr[2][G][4];  // trust-region round vectors: [round parity][rank][column] (PeerLayout tr_off)
  std::atomic<uint64_t> flags[3][G];
  uint64_t epoch[3];  // this rank's own epoch counters: touched by whichever block arrives last
};

struct Rank {
  // local vectors (three rotating buffers where the kernel has them)
  double x[3][N], kty[3][N], y[3][MS], kx[3][MS];
  State state[2];
  double block_partials[B][4];
  double slice_partials[MS / kRowSlice][2];
  std::atomic<unsigned> sync_ticket{0}, sync_gen{0};
  std::atomic<unsigned> tickets[3];  // D, T1, T
  std::atomic<int> error{0};
};

double K[M][N];
Arena arena[G];
Rank rank_[G];

double clampd(double v, double lo, double hi) { return std::fmax(std::fmin(v, hi), lo); }

// ---- the barriers ---------------------------------------------------------------------------
void peer_barrier(int g, int which) {
  Arena& local = arena[g];
  const uint64_t e = local.epoch[which] + 1;
  for (int h = 0; h < G; ++h) arena[h].flags[which][g].store(e, std::memory_order_release);
  for (int h = 0; h < G; ++h) {
    long spins = 0;
    while (local.flags[which][h].load(std::memory_order_acquire) < e) {
      std::this_thread::yield();
      if (++spins > 200000000L) { rank_[g].error.store(1); return; }
    }
  }
  local.epoch[which] = e;
}

enum { kPlain, kPeer, kSumsPeer };
// Grid barrier number `bar` of the rank (loop_grid_sync of the kernel): the block that draws the last ticket
// adds the partials in a fixed order when asked to, does the cross-rank part, then releases the others.
void grid_sync(int g, int blk, unsigned& bar, int kind, int which, bool slice_sums) {
  Rank& r = rank_[g];
  const unsigned ticket = r.sync_ticket.fetch_add(1u, std::memory_order_acq_rel);
  if (ticket == B * (bar + 1u) - 1u) {
    if (kind == kSumsPeer) {
      double t[3] = {0.0, 0.0, 0.0};
      for (int k = 0; k < B; ++k) {
        t[0] += r.block_partials[k][0];
        if (!slice_sums) { t[1] += r.block_partials[k][1]; t[2] += r.block_partials[k][2]; }
      }
      if (slice_sums)
        for (int k = 0; k < MS / kRowSlice; ++k) { t[1] += r.slice_partials[k][0]; t[2] += r.slice_partials[k][1]; }
      for (int h = 0; h < G; ++h)
        for (int k = 0; k < 3; ++k) arena[h].scal[4 * g + k] = t[k];
    }
    bool drop = false;
#ifdef DROP_PEER_A
    drop = drop || which == 0;
#endif
#ifdef DROP_PEER_B
    drop = drop || which == 1;
#endif
    if (kind != kPlain && !drop) peer_barrier(g, which);
    r.sync_gen.store(bar + 1u, std::memory_order_release);
  } else {
    long spins = 0;
    while (r.sync_gen.load(std::memory_order_acquire) < bar + 1u) {
      std::this_thread::yield();
      if (++spins > 400000000L) { r.error.store(1); break; }
    }
  }
  (void)blk;
  ++bar;
}

// ---- the arithmetic of one attempt, shared by the model and the sequential reference -----------
inline void primal_elem(const double* xc, const double* kty, double step, int i, double* xn, double* xt_out, double* dx2) {
  const double nx = clampd(xc[i] - step * (0.01 * i - kty[i]), -3.0, 3.0);
  const double d = nx - xc[i];
  xn[i] = nx;
  *xt_out = nx + d;
  *dx2 += d * d;
}
inline void dual_row(int grow, const double* xt, double yc, double kxc, double step, double* yn, double* kxn, double* dy2, double* dot) {
  double acc = 0.0;
  for (int c = 0; c < N; ++c) acc += K[grow][c] * xt[c];
  const double ny = clampd(yc - step * (0.5 * (grow % 3) - acc), -2.0, 2.0);
  const double knew = 0.5 * (acc + kxc);
  *yn = ny;
  *kxn = knew;
  *dy2 += (ny - yc) * (ny - yc);
  *dot += (knew - kxc) * (ny - yc);
}
inline void decide(State* out, const State& in, const double* scal) {
  double dx2 = 0.0, dy2 = 0.0, dot = 0.0;
  for (int k = 0; k < G; ++k) { dx2 += scal[4 * k]; dy2 += scal[4 * k + 1]; dot += scal[4 * k + 2]; }
  State s = in;
  const bool accept = std::fabs(dot) * in.step <= 0.5 * (dx2 + dy2) + 1e-300;  // a data-dependent, rank-independent choice
  if (accept) { s.cur = in.cand; s.cand = (in.cand + 1) % 3; s.step = in.step * 1.02; } else { s.step = in.step * 0.7; }
  *out = s;
}

// ---- one block of one rank ---------------------------------------------------------------------
void block_main(int g, int blk) {
  Rank& r = rank_[g];
  Arena& mine = arena[g];
  unsigned bar = 0;
  for (int it = 0; it < A; ++it) {
    const State st = r.state[it & 1];  // (the block's copy of the attempt's state)
    State* st_out = &r.state[(it + 1) & 1];
    if (st.halt != 0 || r.error.load() != 0) break;
    const int cur = st.cur, cand = st.cand;
    // P: static element -> block map over the rank's slice
    {
      double s = 0.0;
      for (int i = g * NS + blk; i < (g + 1) * NS; i += B) {
        double xt;
        primal_elem(r.x[cur], r.kty[cur], st.step, i, r.x[cand], &xt, &s);
        for (int h = 0; h < G; ++h) arena[h].xt[i] = xt;
      }
      r.block_partials[blk][0] = s;
    }
    grid_sync(g, blk, bar, kPeer, 0, false);
    // D: row slices by tickets
    if (blk == 0) { r.tickets[1].store(0u, std::memory_order_relaxed); r.tickets[2].store(0u, std::memory_order_relaxed); }
    for (unsigned sl = r.tickets[0].fetch_add(1u, std::memory_order_relaxed); sl < MS / kRowSlice; sl = r.tickets[0].fetch_add(1u, std::memory_order_relaxed)) {
      double red[2] = {0.0, 0.0};
      for (int k = 0; k < kRowSlice; ++k) {
        const int lr = sl * kRowSlice + k, grow = g * MS + lr;
        dual_row(grow, mine.xt, r.y[cur][lr], r.kx[cur][lr], st.step, &r.y[cand][lr], &r.kx[cand][lr], &red[0], &red[1]);
        if (MODE == 0)
          for (int h = 0; h < G; ++h) arena[h].y[grow] = r.y[cand][lr];
      }
      r.slice_partials[sl][0] = red[0];
      r.slice_partials[sl][1] = red[1];
    }
    if (MODE == 1) {
      grid_sync(g, blk, bar, kPlain, 0, false);
      if (blk == 0) r.tickets[0].store(0u, std::memory_order_relaxed);
      const int nsl = MS / kRowSlice, per = (nsl + B - 1) / B, sb = std::min(nsl, per * blk), se = std::min(nsl, sb + per);
      double red[2] = {0.0, 0.0};
      for (int k = sb; k < se; ++k) { red[0] += r.slice_partials[k][0]; red[1] += r.slice_partials[k][1]; }
      r.block_partials[blk][1] = red[0];
      r.block_partials[blk][2] = red[1];
      // T1: the local partial into the own arena, column slices by tickets
      for (unsigned sl = r.tickets[1].fetch_add(1u, std::memory_order_relaxed); sl < N / kColSlice; sl = r.tickets[1].fetch_add(1u, std::memory_order_relaxed))
        for (int c = sl * kColSlice; c < (int)(sl + 1) * kColSlice; ++c) {
          double acc = 0.0;
          for (int lr = 0; lr < MS; ++lr) acc += K[g * MS + lr][c] * r.y[cand][lr];
          mine.partial[c] = acc;
        }
    }
    grid_sync(g, blk, bar, kSumsPeer, 1, MODE == 0);
    if (MODE == 0 && blk == 0) r.tickets[0].store(0u, std::memory_order_relaxed);
    if (blk == B - 1) decide(st_out, st, mine.scal);
    // T
    if (MODE == 0) {
      for (unsigned sl = r.tickets[2].fetch_add(1u, std::memory_order_relaxed); sl < NS / kColSlice; sl = r.tickets[2].fetch_add(1u, std::memory_order_relaxed))
        for (int c = g * NS + sl * kColSlice; c < g * NS + (int)(sl + 1) * kColSlice; ++c) {
          double acc = 0.0;
          for (int row = 0; row < M; ++row) acc += K[row][c] * mine.y[row];
          r.kty[cand][c] = acc;
        }
    } else {
      for (int i = g * NS + blk; i < (g + 1) * NS; i += B) {
        double v = 0.0;
        for (int h = 0; h < G; ++h) v += arena[h].partial[i];
        r.kty[cand][i] = v;
      }
    }
    grid_sync(g, blk, bar, kPlain, 0, false);
  }
}

// ---- the rounds of the trust-region solve (tr_round<PEER> of k_tr_solve) -------------------------
// Every round each rank stores its vector of totals into slot (round & 1) of every arena, the ranks
// meet ONCE, and each adds the G vectors in rank order from its own arena. One barrier per round is
// enough because the slots alternate: a rank can be at most one round ahead of a peer that still
// reads, and then it writes the other slot. -DTR_SINGLE_SLOT is the negative control.
constexpr int kTrRounds = 30;
double tr_result[G][kTrRounds];
void tr_main(int g) {
  double carry = 1.0 + g;
  for (int round = 0; round < kTrRounds; ++round) {
#ifdef TR_SINGLE_SLOT
    const int slot = 0;
#else
    const int slot = round & 1;
#endif
    for (int h = 0; h < G; ++h)
      for (int c = 0; c < 4; ++c) arena[h].tr[slot][g][c] = carry * (c + 1) + round;
    peer_barrier(g, 2);
    double v = 0.0;
    for (int h = 0; h < G; ++h)
      for (int c = 0; c < 4; ++c) v += arena[g].tr[slot][h][c];
    tr_result[g][round] = v;
    carry = 0.5 * carry + 1e-3 * v;  // (the next round's vector depends on this round's totals, as the bracket does)
  }
}

// ---- the same attempts, one thread, no arenas -----------------------------------------------------
void sequential(double x_out[N]) {
  static double x[3][N], kty[3][N], y[3][M], kx[3][M];
  std::memset(x, 0, sizeof x); std::memset(kty, 0, sizeof kty); std::memset(y, 0, sizeof y); std::memset(kx, 0, sizeof kx);
  State st{0, 1, 0, 0.05};
  for (int it = 0; it < A; ++it) {
    double xt[N], scal[4 * G];
    for (int g = 0; g < G; ++g) {
      double part[B];
      for (int blk = 0; blk < B; ++blk) {
        part[blk] = 0.0;
        for (int i = g * NS + blk; i < (g + 1) * NS; i += B) primal_elem(x[st.cur], kty[st.cur], st.step, i, x[st.cand], &xt[i], &part[blk]);
      }
      double t0 = 0.0;
      for (int blk = 0; blk < B; ++blk) t0 += part[blk];
      scal[4 * g] = t0;
    }
    for (int g = 0; g < G; ++g) {
      double sp[MS / kRowSlice][2];
      for (int sl = 0; sl < MS / kRowSlice; ++sl) {
        sp[sl][0] = sp[sl][1] = 0.0;
        for (int k = 0; k < kRowSlice; ++k) {
          const int grow = g * MS + sl * kRowSlice + k;
          dual_row(grow, xt, y[st.cur][grow], kx[st.cur][grow], st.step, &y[st.cand][grow], &kx[st.cand][grow], &sp[sl][0], &sp[sl][1]);
        }
      }
      double t1 = 0.0, t2 = 0.0;
      if (MODE == 0) {
        for (int sl = 0; sl < MS / kRowSlice; ++sl) { t1 += sp[sl][0]; t2 += sp[sl][1]; }
      } else {  // per block over fixed ranges, then the blocks in order
        const int nsl = MS / kRowSlice, per = (nsl + B - 1) / B;
        for (int blk = 0; blk < B; ++blk) {
          const int sb = std::min(nsl, per * blk), se = std::min(nsl, sb + per);
          double a = 0.0, b = 0.0;
          for (int k = sb; k < se; ++k) { a += sp[k][0]; b += sp[k][1]; }
          t1 += a; t2 += b;
        }
      }
      scal[4 * g + 1] = t1;
      scal[4 * g + 2] = t2;
    }
    for (int c = 0; c < N; ++c) {
      double v = 0.0;
      if (MODE == 0) {
        for (int row = 0; row < M; ++row) v += K[row][c] * y[st.cand][row];
      } else {
        for (int h = 0; h < G; ++h) {
          double acc = 0.0;
          for (int lr = 0; lr < MS; ++lr) acc += K[h * MS + lr][c] * y[st.cand][h * MS + lr];
          v += acc;
        }
      }
      kty[st.cand][c] = v;
    }
    State next;
    decide(&next, st, scal);
    st = next;
  }
  std::memcpy(x_out, x[st.cur], sizeof(double) * N);
}
}  // namespace

int main() {
  uint64_t seed = 88172645463325252ull;
  for (int r = 0; r < M; ++r)
    for (int c = 0; c < N; ++c) {
      seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17;
      K[r][c] = (seed % 5 == 0) ? (static_cast<double>(seed % 2001) / 1000.0 - 1.0) : 0.0;
    }
  for (int g = 0; g < G; ++g) {
    std::memset(arena[g].xt, 0, sizeof arena[g].xt); std::memset(arena[g].y, 0, sizeof arena[g].y);
    std::memset(arena[g].partial, 0, sizeof arena[g].partial); std::memset(arena[g].scal, 0, sizeof arena[g].scal);
    for (int w = 0; w < 3; ++w) { arena[g].epoch[w] = 0; for (int h = 0; h < G; ++h) arena[g].flags[w][h].store(0); }
    std::memset(arena[g].tr, 0, sizeof arena[g].tr);
    Rank& r = rank_[g];
    std::memset(r.x, 0, sizeof r.x); std::memset(r.kty, 0, sizeof r.kty); std::memset(r.y, 0, sizeof r.y); std::memset(r.kx, 0, sizeof r.kx);
    r.state[0] = State{0, 1, 0, 0.05};
    r.state[1] = State{0, 1, 0, 0.05};
    for (auto& t : r.tickets) t.store(0u);
  }
  std::vector<std::thread> threads;
  for (int g = 0; g < G; ++g)
    for (int blk = 0; blk < B; ++blk) threads.emplace_back(block_main, g, blk);
  for (auto& t : threads) t.join();
  threads.clear();
  for (int g = 0; g < G; ++g) threads.emplace_back(tr_main, g);
  for (auto& t : threads) t.join();
  double want[N];
  sequential(want);
  int bad = 0;
  for (int g = 0; g < G; ++g) {
    if (rank_[g].error.load() != 0) { std::printf("rank %d: a barrier timed out\n", g); bad = 1; }
    const State& st = rank_[g].state[A & 1];
    for (int i = g * NS; i < (g + 1) * NS; ++i)
      if (std::memcmp(&rank_[g].x[st.cur][i], &want[i], sizeof(double)) != 0) {
        std::printf("rank %d x[%d] = %.17g, sequential %.17g\n", g, i, rank_[g].x[st.cur][i], want[i]);
        bad = 1;
      }
  }
  {  // the trust-region rounds: every rank computed the same totals, equal to the sequential ones
    double carry[G];
    for (int g = 0; g < G; ++g) carry[g] = 1.0 + g;
    for (int round = 0; round < kTrRounds; ++round) {
      double v = 0.0;
      for (int h = 0; h < G; ++h)
        for (int c = 0; c < 4; ++c) v += carry[h] * (c + 1) + round;
      for (int g = 0; g < G; ++g) {
        if (std::memcmp(&tr_result[g][round], &v, sizeof(double)) != 0) {
          std::printf("rank %d trust-region round %d: %.17g, sequential %.17g\n", g, round, tr_result[g][round], v);
          bad = 1;
        }
        carry[g] = 0.5 * carry[g] + 1e-3 * v;
      }
    }
  }
  std::printf(bad ? "MISMATCH\n" : "model ok: mode %d, %d ranks x %d blocks, %d attempts, bitwise equal to the sequential run\n", MODE, G, B, A);
  return bad;
}
