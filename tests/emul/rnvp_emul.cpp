// TEST INFRASTRUCTURE ONLY -- host emulator of the per-tile op programs.
//
// The build container has no GPU, so this executes the *same* program the planner
// (probaforms_b200/csrc/rnvp_planner.h) hands to the CUDA tile kernel, op by op, with plain scalar
// loops over an emulated shared-memory array that starts out as NaN (any read of a location the
// program never wrote poisons the result).  It checks the planner -- offsets, strides, chunking,
// stash, flags, packed layout maps -- against the oracle before GPU time is spent; it says nothing
// about the device micro-kernels, which only the -m gpu tests cover.  Never loaded by the product.
#include <math.h>
#include <stdio.h>
#include <limits>
#include <vector>

#include "../../probaforms_b200/csrc/rnvp_planner.h"

using namespace rnvp_planner;

namespace {

struct Emu {
  const FlowGeom* g;
  Builder* b;
  std::vector<float> sm, packed, gpacked, stash, slot;
  const float *X, *C;
  const long long* idx;
  long long N;
  float scale;
  float *out_x, *out_logdet, *out_logp;
  double loss_sum = 0;
  int mode, R;

  float act(float v, int a) const { return a == 1 ? tanhf(v) : (a == 2 ? fmaxf(v, 0.f) : v); }
  float actp(float h, int a) const { return a == 1 ? 1.f - h * h : (a == 2 ? (h > 0.f ? 1.f : 0.f) : 1.f); }

  void load_chunk(const RnvpChunk& c) {
    slot.assign(b->slot_floats, std::numeric_limits<float>::quiet_NaN());
    const int wn = c.rows_p * c.Ks;
    for (int net = 0; net < 2; ++net) {
      for (int i = 0; i < wn; ++i) slot[net * wn + i] = packed[c.w_src[net] + i];
      for (int i = 0; i < c.rows_p; ++i) slot[2 * wn + net * c.rows_p + i] = packed[c.b_src[net] + i];
    }
  }

  void run_tile(long long row0) {
    const RnvpSmem& s = b->sm;
    const int D = g->D, Cd = g->Cd;
    for (const RnvpOp& op : b->ops) {
      if (op.chunk >= 0) load_chunk(b->chunks[op.chunk]);
      switch (op.kind) {
        case OP_LOAD:
          for (int r = 0; r < R; ++r) {
            const long long row = row0 + r;
            const long long src = row < N ? (idx ? idx[row] : row) : -1;
            for (int j = 0; j < D; ++j) sm[s.xs + r * s.xs_stride + j] = src >= 0 ? X[src * D + j] : 0.f;
            for (int j = 0; j < Cd; ++j) sm[s.cs + r * s.cs_stride + j] = src >= 0 ? C[src * Cd + j] : 0.f;
            sm[s.ld + r] = 0.f;
          }
          break;
        case OP_BUILD_U:
          for (int r = 0; r < R; ++r) {
            if (op.flags & F_ADDGU)
              for (int kk = 0; kk < op.nT; ++kk)
                sm[s.gx + r * s.xs_stride + 2 * kk + op.par] +=
                    sm[s.gu + r * s.gu_stride + kk] + sm[s.gu + s.gu_net + r * s.gu_stride + kk];
            if (op.flags & F_RESTORE)
              for (int ii = 0; ii < op.nT; ++ii)
                sm[s.xs + r * s.xs_stride + 2 * ii + op.par] = stash[op.stash_off + r * op.nT + ii];
            for (int kk = 0; kk < op.Kc; ++kk) {
              float v = 0.f;
              if (kk < op.nK) v = sm[s.xs + r * s.xs_stride + 2 * kk + (1 - op.par)];
              else if (kk < op.nK + Cd) v = sm[s.cs + r * s.cs_stride + kk - op.nK];
              sm[s.ub + r * s.ub_stride + kk] = v;
            }
          }
          break;
        case OP_LINEAR:
          for (int net = 0; net < 2; ++net) {
            if ((op.flags & F_NET_S_ONLY) && net == 0) continue;
            const float* W = slot.data() + net * op.rows_p * op.Ks;
            const float* B = slot.data() + 2 * op.rows_p * op.Ks + net * op.rows_p;
            for (int r = 0; r < R; ++r)
              for (int n = 0; n < op.rows_p; ++n) {
                float acc = 0.f;
                for (int k = 0; k < op.Kc; ++k) acc += sm[op.a_off + net * op.a_net + r * op.a_stride + k] * W[n * op.Ks + k];
                sm[op.o_off + net * op.o_net + r * op.o_stride + op.n0 + n] = act(acc + B[n], op.act);
              }
          }
          break;
        case OP_COUPLE_F:
          for (int r = 0; r < R; ++r)
            for (int ii = 0; ii < op.nT; ++ii) {
              float& x = sm[s.xs + r * s.xs_stride + 2 * ii + op.par];
              const float t = sm[s.st + r * s.st_stride + ii], sv = sm[s.st + s.st_net + r * s.st_stride + ii];
              if (op.flags & F_STASH) stash[op.stash_off + r * op.nT + ii] = x;
              x = x * expf(sv) + t;
              sm[s.ld + r] += sv;
            }
          break;
        case OP_COUPLE_G:
          for (int r = 0; r < R; ++r)
            for (int ii = 0; ii < op.nT; ++ii) {
              float& x = sm[s.xs + r * s.xs_stride + 2 * ii + op.par];
              const float t = sm[s.st + r * s.st_stride + ii], sv = sm[s.st + s.st_net + r * s.st_stride + ii];
              x = (x - t) * expf(-sv);
            }
          break;
        case OP_STORE_F:
          for (int r = 0; r < R; ++r) {
            const bool valid = row0 + r < N;
            float q = 0.f;
            for (int j = 0; j < D; ++j) { const float z = sm[s.xs + r * s.xs_stride + j]; q += z * z; }
            const float ld = sm[s.ld + r];
            const float lp = ld - 0.5f * (D * 1.8378770664093453f + q);
            if (valid) {
              if (mode == 0 && out_x) for (int j = 0; j < D; ++j) out_x[(row0 + r) * D + j] = sm[s.xs + r * s.xs_stride + j];
              if (out_logdet) out_logdet[row0 + r] = ld;
              if (out_logp) out_logp[row0 + r] = lp;
              loss_sum += lp;
            }
            if (mode == 2) {
              sm[s.ld + r] = valid ? scale : 0.f;
              for (int j = 0; j < D; ++j) sm[s.gx + r * s.xs_stride + j] = valid ? -scale * sm[s.xs + r * s.xs_stride + j] : 0.f;
            }
          }
          break;
        case OP_STORE_G:
          for (int r = 0; r < R; ++r)
            if (row0 + r < N) for (int j = 0; j < D; ++j) out_x[(row0 + r) * D + j] = sm[s.xs + r * s.xs_stride + j];
          break;
        case OP_COUPLE_B: {
          const int nTp = (op.nT + 3) & ~3;
          for (int r = 0; r < R; ++r)
            for (int ii = 0; ii < nTp; ++ii) {
              float dt = 0.f, ds = 0.f;
              if (ii < op.nT) {
                float& gy = sm[s.gx + r * s.xs_stride + 2 * ii + op.par];
                const float x = sm[s.xs + r * s.xs_stride + 2 * ii + op.par];
                const float es = expf(sm[s.st + s.st_net + r * s.st_stride + ii]);
                dt = gy; ds = gy * x * es + sm[s.ld + r]; gy = gy * es;
              }
              sm[s.st + r * s.st_stride + ii] = dt;
              sm[s.st + s.st_net + r * s.st_stride + ii] = ds;
            }
        } break;
        case OP_WGRAD:
          for (int net = 0; net < 2; ++net)
            for (int n = 0; n < op.rows_p; ++n) {
              float bsum = 0.f;
              for (int r = 0; r < R; ++r) bsum += sm[op.a_off + net * op.a_net + r * op.a_stride + n];
              gpacked[op.g_b[net] + n] += bsum;
              for (int k = 0; k < op.Kc; ++k) {
                float acc = 0.f;
                for (int r = 0; r < R; ++r)
                  acc += sm[op.a_off + net * op.a_net + r * op.a_stride + n] * sm[op.h_off + net * op.h_net + r * op.h_stride + k];
                gpacked[op.g_w[net] + n * op.Ks + k] += acc;
              }
            }
          break;
        case OP_DGRAD: {
          const int kc = (op.kout + 3) & ~3;
          for (int net = 0; net < 2; ++net) {
            const float* W = slot.data() + net * op.rows_p * op.Ks;
            for (int r = 0; r < R; ++r)
              for (int k = 0; k < kc; ++k) {
                float v = 0.f;
                for (int n = 0; n < op.rows_p; ++n) v += sm[op.a_off + net * op.a_net + r * op.a_stride + op.n0 + n] * W[n * op.Ks + k];
                if (!(op.flags & F_FIRST)) v += sm[op.d_off + net * op.d_net + r * op.d_stride + k];
                if (op.flags & F_LAST) {
                  if (op.flags & F_TO_GU) sm[op.o_off + net * op.o_net + r * op.o_stride + k] = v;
                  else {
                    float& h = sm[op.h_off + net * op.h_net + r * op.h_stride + k];
                    h = v * actp(h, op.act);
                  }
                } else sm[op.d_off + net * op.d_net + r * op.d_stride + k] = v;
              }
          }
        } break;
        default: break;
      }
    }
  }
};

}  // namespace

extern "C" int rnvp_emulate(int mode, int D, int Cd, int L, int nh, const int* hidden, int act, int TR_force,
                            const float* flat, const float* X, const float* C, const long long* idx, long long N,
                            int l0, int l1, float scale, float* out_x, float* out_logdet, float* out_logp,
                            float* gflat, double* loss_sum, int* info /* [TR, smem_bytes, n_ops, n_chunks, packed] */) {
  FlowGeom g;
  g.D = D; g.Cd = Cd; g.L = L; g.nh = nh; g.act = act;
  for (int q = 0; q < nh; ++q) g.hidden[q] = hidden[q];
  build_layout(&g);
  std::vector<int> p2f, f2p;
  build_maps(&g, p2f, f2p);
  Builder b;
  b.d = &g; b.mode = mode; b.l0 = l0; b.l1 = l1;
  const bool ok = b.plan_best(TR_force);
  if (!ok) return -2;
  b.build();
  if (info) { info[0] = b.TR; info[1] = b.sm.total_floats * 4; info[2] = (int)b.ops.size(); info[3] = (int)b.chunks.size(); info[4] = (int)g.packed; }
  Emu e;
  e.g = &g; e.b = &b; e.mode = mode; e.R = b.R;
  e.X = X; e.C = C; e.idx = idx; e.N = N; e.scale = scale;
  e.out_x = out_x; e.out_logdet = out_logdet; e.out_logp = out_logp;
  const float nan = std::numeric_limits<float>::quiet_NaN();
  e.sm.assign(b.sm.total_floats, nan);
  e.stash.assign(b.stash_per_cta, nan);
  e.packed.assign(g.packed_gather, 0.f);
  for (int64_t p = 0; p < g.packed_gather; ++p) e.packed[p] = p2f[p] >= 0 ? flat[p2f[p]] : 0.f;
  e.gpacked.assign(g.packed_tile, 0.f);
  for (long long row0 = 0; row0 < N; row0 += b.R) e.run_tile(row0);
  if (gflat) for (int64_t f = 0; f < g.P; ++f) gflat[f] = f2p[f] >= 0 ? e.gpacked[f2p[f]] : 0.f;
  if (loss_sum) *loss_sum = e.loss_sum;
  return 0;
}

extern "C" long long rnvp_emul_param_count(int D, int Cd, int L, int nh, const int* hidden) {
  FlowGeom g;
  g.D = D; g.Cd = Cd; g.L = L; g.nh = nh; g.act = 1;
  for (int q = 0; q < nh; ++q) g.hidden[q] = hidden[q];
  build_layout(&g);
  return g.P;
}
