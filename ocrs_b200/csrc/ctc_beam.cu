// CTC prefix beam search on the GPU: one thread block per text line.
//
// Stands where `CtcDecoder::decode_beam(seq, width)` is called in the reference
// (ocrs/src/recognition.rs:512-514; `DecodeMethod::BeamSearch { width }` :199-205, CLI width 100
// ocrs-cli/src/main.rs:403-404).  The decoder itself lives in the un-vendored rten-text crate, so
// the algorithm restated here is the published one (Hannun et al. 2014, prefix beam search without
// a language model, log space) and is checked against oracle/recognition.py::ctc_decode_beam.
//
// Per line the block keeps `width` prefixes as nodes of a trie (parent, label, timestep) in
// global memory plus (log p_blank, log p_non_blank) per prefix in shared memory.  A prefix can be
// created more than once (it drops out of the beam and is reached again later, with new timesteps), so
// every node also carries the CANONICAL id of its prefix -- the id of the first node ever created for
// (canonical parent, label), found through per-node child lists -- and prefix identity is decided on
// canonical ids, exactly like a decoder that keys its beam on the label tuple.  Every timestep:
//   1. each beam's "stay" entry is scored (blank path + repeat path + the extension of its
//      parent prefix when the parent prefix is also in the beam -- the only way two candidates can
//      name the same prefix);
//   2. every (beam, class) extension that is not merged into a stay entry is a candidate;
//   3. the `width` best candidates (score descending, creation order ascending, i.e. the order in which
//      a sequential implementation walking beams by rank and classes by index first touches
//      each prefix) are selected
//      exactly with a 64-bit radix select over the candidate scores, then bitonic-sorted.
// All arithmetic is double precision; this TU is compiled with -fmad=false.
#include <cfloat>
#include <climits>

#include <string>

#include "common.h"
#include "image_kernels.h"

namespace ocrs {
namespace img {
namespace {

constexpr int kBeamThreads = 128;
constexpr int kNodeInts = 6;
constexpr int kHistBins = 2048;

__device__ __forceinline__ double lae(double a, double b) {
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  double m = fmax(a, b);
  return m + log1p(exp(-fabs(a - b)));
}

__device__ __forceinline__ unsigned long long order_key(double v) {
  unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

struct BeamShared {
  double* lp;      // [C]
  double *b_pb, *b_pnb, *b_tot;  // current beams
  double *s_pb, *s_pnb, *s_tot;  // their stay entries at this timestep
  double *n_pb, *n_pnb;          // next beams
  double* sel_score;             // [Wp2]
  int *b_node, *b_parent, *b_last, *b_pj, *b_oid, *b_canon, *b_pcanon;
  int *n_node, *n_parent, *n_last, *n_canon, *n_pcanon;
  int *sel_id, *sel_oid;         // [Wp2] candidate id, creation-order id
  unsigned* mask;                // [W * Cw]
  int* hist;                     // [kHistBins]
  int* misc;                     // [8]
};

__device__ __forceinline__ double cand_score(const BeamShared& s, int i, int c, int Cw) {
  if (c == 0) return s.s_tot[i];
  if (s.mask[i * Cw + (c >> 5)] >> (c & 31) & 1u) return -INFINITY;  // merged into a stay entry
  double base = (c == s.b_last[i] && s.b_node[i] != 0) ? s.b_pb[i] : s.b_tot[i];
  return base + s.lp[c];
}

// Finds, among the bins of hist[0, nb), the bin d (scanning from the top when `descending`)
// where the running count reaches kk.  Returns d and the residual count in misc[0], misc[1].
__device__ void find_bin(const int* hist, int nb, int kk, bool descending, int* misc) {
  if (threadIdx.x >= 32) return;
  int lane = threadIdx.x;
  int chunk = nb / 32;
  int sum = 0;
  for (int q = 0; q < chunk; ++q) {
    int b = descending ? nb - 1 - (lane * chunk + q) : lane * chunk + q;
    sum += hist[b];
  }
  int incl = sum;
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  unsigned ball = __ballot_sync(0xffffffffu, incl >= kk);
  int first = __ffs(ball) - 1;
  if (lane == first) {
    int before = incl - sum;
    for (int q = 0; q < chunk; ++q) {
      int b = descending ? nb - 1 - (lane * chunk + q) : lane * chunk + q;
      int h = hist[b];
      if (before + h >= kk) {
        misc[0] = b;
        misc[1] = kk - before;
        misc[2] = h;
        break;
      }
      before += h;
    }
  }
}

__global__ void __launch_bounds__(kBeamThreads) ctc_beam_kernel(const float* __restrict__ logits, int C,
                                                                const uint8_t* __restrict__ excluded,
                                                                const CtcLine* __restrict__ lines, int W, int Wp2,
                                                                int32_t* __restrict__ nodes,
                                                                int32_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const CtcLine L = lines[blockIdx.x];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int Cw = (C + 31) / 32;
  BeamShared s;
  {
    double* d = reinterpret_cast<double*>(smem_raw);
    s.lp = d; d += (C + 1) & ~1;
    s.b_pb = d; d += W; s.b_pnb = d; d += W; s.b_tot = d; d += W;
    s.s_pb = d; d += W; s.s_pnb = d; d += W; s.s_tot = d; d += W;
    s.n_pb = d; d += W; s.n_pnb = d; d += W;
    s.sel_score = d; d += Wp2;
    int* q = reinterpret_cast<int*>(d);
    s.b_node = q; q += W; s.b_parent = q; q += W; s.b_last = q; q += W; s.b_pj = q; q += W; s.b_oid = q; q += W;
    s.b_canon = q; q += W; s.b_pcanon = q; q += W;
    s.n_node = q; q += W; s.n_parent = q; q += W; s.n_last = q; q += W; s.n_canon = q; q += W; s.n_pcanon = q; q += W;
    s.sel_id = q; q += Wp2; s.sel_oid = q; q += Wp2;
    s.mask = reinterpret_cast<unsigned*>(q); q += W * Cw;
    s.hist = q; q += kHistBins;
    s.misc = q;
  }
  // node record: [parent, label, pos, canonical id, next sibling, first child]; the sibling lists hang off the
  // CANONICAL node of the parent prefix and hold one node per distinct (parent prefix, label)
  int32_t* nd = nodes + L.node_off * kNodeInts;
  int nb = 1;
  if (tid == 0) {
    s.b_node[0] = 0; s.b_parent[0] = -1; s.b_last[0] = 0; s.b_canon[0] = 0; s.b_pcanon[0] = -1;
    s.b_pb[0] = 0.0; s.b_pnb[0] = -INFINITY; s.b_tot[0] = 0.0;
    nd[0] = -1; nd[1] = 0; nd[2] = 0; nd[3] = 0; nd[4] = -1; nd[5] = -1;
  }
  __syncthreads();

  for (int t = 0; t < L.T; ++t) {
    const float* row = logits + (L.base + (int64_t)t * L.stride) * C;
    for (int c = tid; c < C; c += nthr) s.lp[c] = (excluded && excluded[c]) ? -INFINITY : (double)row[c];
    for (int w = tid; w < nb * Cw; w += nthr) s.mask[w] = 0u;
    // parent-in-beam lookup
    for (int i = tid; i < nb; i += nthr) {
      int par = s.b_pcanon[i], pj = -1;  // prefix identity = canonical id
      if (par >= 0)
        for (int j = 0; j < nb; ++j)
          if (s.b_canon[j] == par) { pj = j; break; }
      s.b_pj[i] = pj;
    }
    __syncthreads();
    // stay entries and merge mask
    for (int i = tid; i < nb; i += nthr) {
      double tot = s.b_tot[i];
      double pb = tot + s.lp[0];
      double pnb = -INFINITY;
      int oid = i * C;  // creation order of this entry among the step's candidates
      if (s.b_node[i] != 0) {
        int l = s.b_last[i];
        pnb = s.b_pnb[i] + s.lp[l];
        int pj = s.b_pj[i];
        if (pj >= 0) {
          double base = (s.b_last[pj] == l && s.b_node[pj] != 0) ? s.b_pb[pj] : s.b_tot[pj];
          pnb = lae(pnb, base + s.lp[l]);
          atomicOr(&s.mask[pj * Cw + (l >> 5)], 1u << (l & 31));
          // a better-ranked parent reaches this prefix first, at its class-l extension
          if (pj < i && s.lp[l] != -INFINITY) oid = pj * C + l;
        }
      }
      s.b_oid[i] = oid;
      s.s_pb[i] = pb;
      s.s_pnb[i] = pnb;
      s.s_tot[i] = lae(pb, pnb);
    }
    __syncthreads();

    const int n_cand = nb * C;
    // number of finite candidates
    if (tid == 0) s.misc[3] = 0;
    __syncthreads();
    {
      int local = 0;
      for (int id = tid; id < n_cand; id += nthr) {
        int i = id / C, c = id - i * C;
        if (cand_score(s, i, c, Cw) != -INFINITY) ++local;
      }
      for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
      if ((tid & 31) == 0 && local) atomicAdd(&s.misc[3], local);
    }
    __syncthreads();
    const int n_valid = s.misc[3];
    const int K = min(W, n_valid);
    if (K == 0) continue;  // nothing can be extended: keep the beams (uniform across the block)

    // ---- radix select of the K-th largest score ----
    unsigned long long prefix = 0;
    int kk = K, n_ties = 0;
    {
      const int shifts[6] = {53, 42, 31, 20, 9, 0};
      const int widths[6] = {11, 11, 11, 11, 11, 9};
#pragma unroll 1
      for (int p = 0; p < 6; ++p) {
        const int sh = shifts[p], wd = widths[p], nbins = 1 << wd;
        for (int b = tid; b < nbins; b += nthr) s.hist[b] = 0;
        __syncthreads();
        for (int id = tid; id < n_cand; id += nthr) {
          int i = id / C, c = id - i * C;
          double v = cand_score(s, i, c, Cw);
          if (v == -INFINITY) continue;
          unsigned long long key = order_key(v);
          if (p == 0 || (key >> (sh + wd)) == prefix) atomicAdd(&s.hist[(int)((key >> sh) & (unsigned)(nbins - 1))], 1);
        }
        __syncthreads();
        find_bin(s.hist, nbins, kk, true, s.misc);
        __syncthreads();
        prefix = (prefix << wd) | (unsigned long long)s.misc[0];
        kk = s.misc[1];
        n_ties = s.misc[2];
        __syncthreads();
      }
    }
    const unsigned long long tau = prefix;
    // ---- ties on the threshold score: the kk smallest candidate ids win ----
    int id_tau = INT_MAX;
    if (n_ties > kk) {
      unsigned idp = 0;
      int k2 = kk;
#pragma unroll 1
      for (int p = 0; p < 2; ++p) {  // ids < 2^22
        const int sh = p == 0 ? 11 : 0;
        for (int b = tid; b < kHistBins; b += nthr) s.hist[b] = 0;
        __syncthreads();
        for (int id = tid; id < n_cand; id += nthr) {
          int i = id / C, c = id - i * C;
          double v = cand_score(s, i, c, Cw);
          if (v == -INFINITY || order_key(v) != tau) continue;
          int oid = c == 0 ? s.b_oid[i] : id;
          if (p == 0 || (unsigned)(oid >> 11) == idp) atomicAdd(&s.hist[(oid >> sh) & (kHistBins - 1)], 1);
        }
        __syncthreads();
        find_bin(s.hist, kHistBins, k2, false, s.misc);
        __syncthreads();
        idp = (idp << (p == 0 ? 0 : 11)) | (unsigned)s.misc[0];
        k2 = s.misc[1];
        __syncthreads();
      }
      id_tau = (int)idp;
    }
    // ---- gather the survivors ----
    if (tid == 0) s.misc[4] = 0;
    for (int r = tid; r < Wp2; r += nthr) { s.sel_score[r] = -INFINITY; s.sel_id[r] = INT_MAX; s.sel_oid[r] = INT_MAX; }
    __syncthreads();
    for (int id = tid; id < n_cand; id += nthr) {
      int i = id / C, c = id - i * C;
      double v = cand_score(s, i, c, Cw);
      if (v == -INFINITY) continue;
      unsigned long long key = order_key(v);
      int oid = c == 0 ? s.b_oid[i] : id;
      if (key > tau || (key == tau && oid <= id_tau)) {
        int slot = atomicAdd(&s.misc[4], 1);
        if (slot < Wp2) { s.sel_score[slot] = v; s.sel_id[slot] = id; s.sel_oid[slot] = oid; }
      }
    }
    __syncthreads();
    // ---- bitonic sort: score descending, id ascending ----
    for (int k = 2; k <= Wp2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int x = tid; x < Wp2; x += nthr) {
          int y = x ^ j;
          if (y > x) {
            double sx = s.sel_score[x], sy = s.sel_score[y];
            int ox = s.sel_oid[x], oy = s.sel_oid[y];
            bool x_first = sx > sy || (sx == sy && ox < oy);
            bool up = (x & k) == 0;
            if (up ? !x_first : x_first) {
              int ix = s.sel_id[x], iy = s.sel_id[y];
              s.sel_score[x] = sy; s.sel_score[y] = sx;
              s.sel_id[x] = iy; s.sel_id[y] = ix;
              s.sel_oid[x] = oy; s.sel_oid[y] = ox;
            }
          }
        }
        __syncthreads();
      }
    }
    // ---- next beams ----
    // phase 1 (read only): new nodes look their prefix up in the child list of the parent's canonical node
    for (int r = tid; r < K; r += nthr) {
      int id = s.sel_id[r];
      int i = id / C, c = id - i * C;
      if (c == 0) {
        s.n_node[r] = s.b_node[i]; s.n_parent[r] = s.b_parent[i]; s.n_last[r] = s.b_last[i];
        s.n_canon[r] = s.b_canon[i]; s.n_pcanon[r] = s.b_pcanon[i];
        s.n_pb[r] = s.s_pb[i]; s.n_pnb[r] = s.s_pnb[i];
      } else {
        int node = 1 + t * W + r;
        int par = s.b_node[i], pc = s.b_canon[i];
        int canon = -1;
        for (int ch = nd[kNodeInts * (int64_t)pc + 5]; ch >= 0; ch = nd[kNodeInts * (int64_t)ch + 4])
          if (nd[kNodeInts * (int64_t)ch + 1] == c) { canon = ch; break; }
        int32_t* rec = nd + kNodeInts * (int64_t)node;
        rec[0] = par; rec[1] = c; rec[2] = t; rec[3] = canon < 0 ? node : canon; rec[4] = -1; rec[5] = -1;
        s.n_node[r] = node; s.n_parent[r] = par; s.n_last[r] = c;
        s.n_canon[r] = canon < 0 ? node : canon; s.n_pcanon[r] = pc;
        s.n_pb[r] = -INFINITY; s.n_pnb[r] = s.sel_score[r];
      }
    }
    __syncthreads();
    // phase 2: first-time prefixes become the canonical node: push onto the parent's child list
    for (int r = tid; r < K; r += nthr) {
      const int node = s.n_node[r];
      if (node == 1 + t * W + r && s.n_canon[r] == node) {
        const int old = atomicExch(&nd[kNodeInts * (int64_t)s.n_pcanon[r] + 5], node);
        nd[kNodeInts * (int64_t)node + 4] = old;
      }
    }
    for (int r = tid; r < K; r += nthr) {
      s.b_node[r] = s.n_node[r]; s.b_parent[r] = s.n_parent[r]; s.b_last[r] = s.n_last[r];
      s.b_canon[r] = s.n_canon[r]; s.b_pcanon[r] = s.n_pcanon[r];
      s.b_pb[r] = s.n_pb[r]; s.b_pnb[r] = s.n_pnb[r]; s.b_tot[r] = s.sel_score[r];
    }
    __threadfence_block();
    nb = K;
    __syncthreads();
  }

  if (tid == 0) {
    __threadfence_block();
    int n = 0;
    for (int node = s.b_node[0]; node != 0; node = nd[kNodeInts * (int64_t)node]) ++n;
    int k = n;
    for (int node = s.b_node[0]; node != 0; node = nd[kNodeInts * (int64_t)node]) {
      --k;
      out[L.lab_off + k] = nd[kNodeInts * (int64_t)node + 1];
      out[L.pos_off + k] = nd[kNodeInts * (int64_t)node + 2];
    }
    out[L.cnt_off] = n;
  }
}

int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

size_t beam_smem_bytes(int C, int W, int Wp2) {
  size_t Cw = (size_t)(C + 31) / 32;
  return (size_t)(((C + 1) & ~1) + 8 * W + Wp2) * 8 + (size_t)(12 * W + 2 * Wp2 + W * Cw + kHistBins + 8) * 4;
}

}  // namespace

int64_t ctc_beam_nodes_per_line(int T, int width) { return (int64_t)T * width + 1; }

void ctc_beam_search(const float* logits, int C, const uint8_t* excluded, const CtcLine* lines, int n_lines, int width,
                     int32_t* nodes, int32_t* out, cudaStream_t st) {
  if (n_lines == 0) return;
  OCRS_CHECK(width >= 1 && width <= kMaxBeamWidth, kInvalidArg,
             "beam width must be in [1, " + std::to_string(kMaxBeamWidth) + "]");
  OCRS_CHECK((int64_t)width * C < (1 << 22), kInvalidArg, "beam width x classes too large");
  int Wp2 = next_pow2(width);
  size_t smem = beam_smem_bytes(C, width, Wp2);
  OCRS_CHECK(smem <= 227 * 1024, kInvalidArg, "beam width too large for shared memory");
  if (smem > 48 * 1024)  // per-device attribute, cheap to repeat
    OCRS_CUDA_CHECK(cudaFuncSetAttribute(ctc_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  ctc_beam_kernel<<<n_lines, kBeamThreads, smem, st>>>(logits, C, excluded, lines, width, Wp2, nodes, out);
  count_launch();
  OCRS_CUDA_CHECK(cudaGetLastError());
}

}  // namespace img
}  // namespace ocrs
