// Host-side check of the MTTKRP work partition (cp-cals_b200/csrc/mttkrp.cuh: mttkrp_make_plan / PlanView) that the
// scheduler builds on the device and the MTTKRP + reduce kernels read.  Compiled with nvcc, runs on the CPU.
#include <cstdio>
#include <vector>

#include "../../cp-cals_b200/csrc/mttkrp.cuh"

using namespace calsb200;

int main() {
  long long checked = 0;
  for (int G : {1, 2, 7, 148})
    for (int In : {1, 8, 41, 100, 200, 299, 1000})
      for (int WM : {4, 5, 6, 7})
        for (int C : {1, 7, 52, 220, 256, 257, 2100, 2691})
          for (int shape_id = 0; shape_id < 7; shape_id++) {
            // (Ip, Iq, S): single stage, ragged K tail, ragged outer run (41 = 5 * 8 + 1), config 2/3/4-like, 5 modes
            static const int shapes[7][3] = {{1, 1, 1},    {7, 9, 1},    {301, 41, 1}, {200, 200, 1},
                                             {299, 301, 1}, {80, 80, 80}, {20, 5, 12}};
            PlanShape sh{In, WM, shapes[shape_id][0], shapes[shape_id][1], shapes[shape_id][2]};
            const int Tp = (int)plan_tp(sh);
            const int In8 = (In + 7) / 8, NO = (C + 63) / 64;
            const int m_tiles = (In8 + WM - 1) / WM, n_tiles = (NO + OCT_TILE - 1) / OCT_TILE;
            const int pairs = m_tiles * n_tiles;
            std::vector<int> plan(plan_capacity(G, pairs) + 4, -12345);
            mttkrp_make_plan(plan.data(), sh, C, G);
            if (plan[plan_capacity(G, pairs)] != -12345) {
              printf("FAIL overrun\n");
              return 1;
            }
            PlanView pv = plan_view(plan.data(), G);
            const long long total = (long long)pairs * Tp;
            if (pv.pairs != pairs || pv.m_tiles != m_tiles || pv.n_tiles != n_tiles || pv.Tp != Tp) {
              printf("FAIL header\n");
              return 1;
            }
            // even m split covers all row groups with tiles of at most WM
            int cover = 0;
            for (int mt = 0; mt < m_tiles; mt++) {
              int w = plan_wm(mt, In8, m_tiles);
              if (w < 1 || w > WM || plan_m8_start(mt, In8, m_tiles) != cover) {
                printf("FAIL msplit In=%d WM=%d mt=%d\n", In, WM, mt);
                return 1;
              }
              cover += w;
            }
            if (cover != In8) {
              printf("FAIL mcover\n");
              return 1;
            }
            // even octet split covers all columns with tiles of at most 4 octets; plan_tile_of inverts both splits
            cover = 0;
            for (int nt = 0; nt < n_tiles; nt++) {
              int w = plan_nn(nt, NO, n_tiles);
              if (w < 1 || w > OCT_TILE || plan_oct_start(nt, NO, n_tiles) != cover) {
                printf("FAIL nsplit C=%d nt=%d\n", C, nt);
                return 1;
              }
              for (int o = cover; o < cover + w; o++)
                if (plan_tile_of(o, NO, n_tiles) != nt) {
                  printf("FAIL tile_of n\n");
                  return 1;
                }
              cover += w;
            }
            if (cover != NO) {
              printf("FAIL ncover\n");
              return 1;
            }
            for (int m8 = 0; m8 < In8; m8++) {
              int mt = plan_tile_of(m8, In8, m_tiles);
              if (plan_m8_start(mt, In8, m_tiles) > m8 || m8 >= plan_m8_start(mt, In8, m_tiles) + plan_wm(mt, In8, m_tiles)) {
                printf("FAIL tile_of m\n");
                return 1;
              }
            }
            // ranges: contiguous, non-empty for b < G_eff, cover [0,total)
            if (pv.cta_lo[0] != 0 || pv.cta_lo[G] != total || pv.G_eff != (total < G ? (int)total : G)) {
              printf("FAIL ends G=%d In=%d C=%d Tp=%d\n", G, In, C, Tp);
              return 1;
            }
            std::vector<int> seg_pair(pv.n_segments, -1);
            long long wmax = 0, wsum = 0, wchunk_max = 0;
            for (int b = 0; b < G; b++) {
              int lo = pv.cta_lo[b], hi = pv.cta_lo[b + 1];
              if (hi < lo || (b < pv.G_eff && hi == lo) || (b >= pv.G_eff && hi != lo)) {
                printf("FAIL range b=%d\n", b);
                return 1;
              }
              // replay the kernel's segment numbering
              int seg = pv.cta_seg0[b] - 1, prev_pair = -1;
              long long wcta = 0;
              for (int ch = lo; ch < hi; ch++) {
                int pair = ch / Tp;
                if (pair != prev_pair) {
                  seg++;
                  prev_pair = pair;
                  if (seg < 0 || seg >= pv.n_segments || seg_pair[seg] != -1) {
                    printf("FAIL seg id G=%d In=%d C=%d Tp=%d b=%d seg=%d\n", G, In, C, Tp, b, seg);
                    return 1;
                  }
                  seg_pair[seg] = pair;
                }
                int nt = pair / m_tiles, mt = pair - nt * m_tiles;
                const PairCost pc = plan_pair_cost(sh, plan_wm(mt, In8, m_tiles), plan_nn(nt, NO, n_tiles),
                                                   nt == n_tiles - 1 ? (C - 64 * (NO - 1) + 7) / 8 : 8);
                // the rest of this CTA's range inside this pair in one step (the cost model has a closed-form prefix)
                const long long k0 = ch - (long long)pair * Tp;
                long long k1 = (long long)hi - (long long)pair * Tp;
                if (k1 > Tp)
                  k1 = Tp;
                wcta += pc.prefix(k1) - pc.prefix(k0);
                const long long one = pc.prefix(1); // one stage in a full K tile (or the only kind there is)
                if (one > wchunk_max)
                  wchunk_max = one;
                // prefix is consistent with a chunk-by-chunk walk
                if (k1 - k0 <= 50) {
                  long long walk = 0;
                  for (long long k = k0; k < k1; k++)
                    walk += pc.prefix(k + 1) - pc.prefix(k);
                  if (walk != pc.prefix(k1) - pc.prefix(k0) || pc.prefix(k0 + 1) <= pc.prefix(k0)) {
                    printf("FAIL prefix\n");
                    return 1;
                  }
                }
                ch = (int)((long long)pair * Tp + k1) - 1;
              }
              wsum += wcta;
              if (wcta > wmax)
                wmax = wcta;
            }
            // every segment belongs to exactly the pair whose table range holds it; pairs' segments are consecutive
            for (int p = 0; p < pairs; p++) {
              if (pv.pair_seg0[p + 1] <= pv.pair_seg0[p]) {
                printf("FAIL pair without segment\n");
                return 1;
              }
              for (int s = pv.pair_seg0[p]; s < pv.pair_seg0[p + 1]; s++)
                if (seg_pair[s] != p) {
                  printf("FAIL seg->pair G=%d In=%d C=%d Tp=%d pair=%d seg=%d has %d\n", G, In, C, Tp, p, s, seg_pair[s]);
                  return 1;
                }
            }
            if (pv.pair_seg0[pairs] != pv.n_segments || pv.n_segments > G + pairs) {
              printf("FAIL nseg\n");
              return 1;
            }
            // balance: the heaviest CTA is within one chunk (+ rounding) of the ideal share, when there is slack
            if (total >= 4LL * G) {
              double ideal = (double)wsum / pv.G_eff;
              if (wmax > ideal + 2.0 * wchunk_max + 1) {
                printf("FAIL balance G=%d In=%d WM=%d C=%d Tp=%d wmax=%lld ideal=%.1f chunk=%lld\n", G, In, WM, C, Tp, wmax,
                       ideal, wchunk_max);
                return 1;
              }
            }
            checked++;
          }
  printf("OK %lld configurations\n", checked);
  return 0;
}
