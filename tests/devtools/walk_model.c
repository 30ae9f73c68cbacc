// walk_model.c — development aid (NOT product, NOT test): replays the warp-level traversal schedule of
// csrc/walk2.cuh on the CPU over the oracle's canonical tree and counts the events that set its instruction count
// (pop batches and their fill, far / near / mixed / leaf nodes, deferred rounds, list entries and their mask
// population).  Used to size queues and to predict the effect of schedule changes without GPU time.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../oracle/kdtree_oracle.h"

typedef struct {
  double batches, popped, farn, nearn, mixed, leaves, entries, lanes, mixed_rounds, mixed_round_nodes, leaf_rounds,
      leaf_round_nodes, max_sp, max_pend, drains, part_entries, tests, pack2, pack4, pack32, hist[33], tl[24], sw[16];
} wm_stats;

static inline int popc(uint32_t x) { return __builtin_popcount(x); }
static int pack(const uint32_t* m, int n, int K) {
  uint32_t occ[64]; int cnt[64]; int ns = 0;
  for (int i = 0; i < n; ++i) {
    if (!m[i]) continue;
    int j = 0;
    for (; j < ns; ++j) if (!(occ[j] & m[i]) && cnt[j] < K) break;
    if (j == ns) occ[ns] = 0, cnt[ns] = 0, ns++;
    occ[j] |= m[i]; cnt[j]++;
  }
  return ns;
}
static void packstats(wm_stats* o, const uint32_t* m, int n, int rep) {
  o->pack2 += rep * pack(m, n, 2); o->pack4 += rep * pack(m, n, 4); o->pack32 += rep * pack(m, n, 32);
  for (int i = 0; i < n; ++i) if (m[i]) o->hist[popc(m[i])] += rep;
}


// ---- two-list drain model: entries with popc(mask) >= T go to a dense list drained by the broadcast loop (one
// iteration per entry), the others to a sparse list drained lane-privately (rounds = max over lanes of the lane's
// items in the batch).  tl[3*c+0] = dense iterations, tl[3*c+1] = sparse rounds, tl[3*c+2] = sparse drains, for the
// (T, sparse capacity) configurations below.
static const int TL_T[8] = {33, 24, 28, 30, 28, 28, 32, 28};
static const int TL_CAP[8] = {96, 96, 96, 96, 64, 128, 96, 192};
typedef struct { int nd, ns; int col[32]; } tl_state;
static tl_state g_tl[8];
static void tl_flush_sparse(wm_stats* o, int c) {
  tl_state* t = &g_tl[c];
  if (!t->ns) return;
  int mx = 0;
  for (int l = 0; l < 32; ++l) { if (t->col[l] > mx) mx = t->col[l]; t->col[l] = 0; }
  o->tl[3 * c + 1] += mx;
  o->tl[3 * c + 2] += 1;
  t->ns = 0;
}
static void sw_emit(wm_stats* o, uint32_t m);
// ---- tile-drain model: the list (capacity TI_CAP[c]) is drained once per sub-group of g targets (g = 16, 8, 4) with
// 32/g lanes per target, U interactions in flight per lane: a pass over a sub-group costs ceil(cnt_sub / (L*U)) * U
// lane-iterations, cnt_sub = entries whose mask meets the sub-group.  ti[c][0] = today's iterations (ceil(cnt/4)*4),
// ti[c][1..] = (g=16,U=4) (g=8,U=4) (g=8,U=2) (g=4,U=2) (g=4,U=1) (g=8,U=1) ; ti[c][7] = drains
static const int TI_CAP[4] = {96, 128, 192, 256};
double g_ti[4][8];
typedef struct { int cnt; uint32_t m[256]; } ti_state;
static ti_state g_tis[4];
static int ti_pass(const ti_state* t, int g, int U) {
  const int L = 32 / g; int tot = 0;
  for (int sub = 0; sub < 32 / g; ++sub) {
    const uint32_t sm = (g == 32 ? 0xffffffffu : ((1u << g) - 1u)) << (sub * g);
    int c = 0;
    for (int i = 0; i < t->cnt; ++i) c += (t->m[i] & sm) != 0;
    tot += (c + L * U - 1) / (L * U) * U;
  }
  return tot;
}
static void ti_flush(int c) {
  ti_state* t = &g_tis[c];
  if (!t->cnt) return;
  g_ti[c][0] += (t->cnt + 3) / 4 * 4;
  g_ti[c][1] += ti_pass(t, 16, 4);
  g_ti[c][2] += ti_pass(t, 8, 4);
  g_ti[c][3] += ti_pass(t, 8, 2);
  g_ti[c][4] += ti_pass(t, 4, 2);
  g_ti[c][5] += ti_pass(t, 4, 1);
  g_ti[c][6] += ti_pass(t, 8, 1);
  g_ti[c][7] += 1;
  t->cnt = 0;
}
static void ti_emit(uint32_t m) {
  for (int c = 0; c < 4; ++c) {
    if (g_tis[c].cnt == TI_CAP[c]) ti_flush(c);
    g_tis[c].m[g_tis[c].cnt++] = m;
  }
}
static void ti_end_group(void) { for (int c = 0; c < 4; ++c) ti_flush(c); }
static void tl_emit(wm_stats* o, uint32_t m) {
  if (!m) return;
  sw_emit(o, m);
  ti_emit(m);
  for (int c = 0; c < 8; ++c) {
    tl_state* t = &g_tl[c];
    if (popc(m) >= TL_T[c]) {
      o->tl[3 * c + 0] += 1;
    } else {
      if (t->ns == TL_CAP[c]) tl_flush_sparse(o, c);
      t->ns++;
      for (int l = 0; l < 32; ++l) t->col[l] += (m >> l) & 1;
    }
  }
}
static void tl_end_group(wm_stats* o) { for (int c = 0; c < 8; ++c) tl_flush_sparse(o, c); }


// ---- sliding-window model: entries with popc < T enter a ring of W entries; every lane keeps the bit column of its
// pending items; a round lets every lane with a non-empty column consume its OLDEST item.  Rounds run only when the
// ring needs room (until the oldest entries are consumed by every lane) and at the end of the group.
// sw[2*c] = dense iterations, sw[2*c+1] = rounds.
static const int SW_T[8] = {33, 33, 33, 28, 28, 28, 24, 33};
static const int SW_W[8] = {64, 96, 128, 64, 96, 128, 64, 256};
typedef struct { int head, cnt; uint32_t m[256]; } sw_state;  // m[slot] = lanes that still have to consume it
static sw_state g_sw[8];
static void sw_round(wm_stats* o, int c) {
  sw_state* t = &g_sw[c];
  const int W = SW_W[c];
  uint32_t done = 0;  // lanes that consumed an item this round
  for (int k = 0; k < t->cnt && done != 0xffffffffu; ++k) {
    const int s = (t->head + k) % W;
    const uint32_t take = t->m[s] & ~done;
    t->m[s] &= ~take;
    done |= take;
  }
  o->sw[2 * c + 1] += 1;
  while (t->cnt && !t->m[t->head]) t->head = (t->head + 1) % W, t->cnt--;
}
static void sw_emit(wm_stats* o, uint32_t m) {
  if (!m) return;
  for (int c = 0; c < 8; ++c) {
    sw_state* t = &g_sw[c];
    if (popc(m) >= SW_T[c]) { o->sw[2 * c] += 1; continue; }
    while (t->cnt == SW_W[c]) sw_round(o, c);
    t->m[(t->head + t->cnt) % SW_W[c]] = m;
    t->cnt++;
  }
}
static void sw_end_group(wm_stats* o) {
  for (int c = 0; c < 8; ++c) { while (g_sw[c].cnt) sw_round(o, c); g_sw[c].head = 0; }
}

// group = 32 consecutive tree slots starting at base; slot -> particle id through idx[]
// optional explicit groups (wm_set_groups): group g = slots [gbase[g], gbase[g] + gcount[g]), gcount <= 32
static const uint64_t* g_gbase; static const uint32_t* g_gcount;
void wm_set_groups(const uint64_t* b, const uint32_t* c) { g_gbase = b; g_gcount = c; }
// optional lane -> slot map per group (32 entries each, ~0 = empty lane); overrides base + lane
static const uint64_t* g_lanes;
void wm_set_lanes(const uint64_t* m) { g_lanes = m; }
#define SLOT(g_, l_) (g_lanes ? g_lanes[(g_) * 32 + (l_)] : (base + (l_) < gend ? base + (l_) : ~0ull))
void wm_run(const okd_node* nodes, const okd_particle* parts, const uint64_t* idx, uint64_t n, double theta,
            uint64_t first_group, uint64_t n_groups, uint64_t stride, int defer, int hard, wm_stats* out) {
  const double theta2 = theta * theta;
  memset(out, 0, sizeof(*out));
  memset(g_ti, 0, sizeof(g_ti));
  enum { CAP = 4096 };
  uint32_t* snode = malloc(sizeof(uint32_t) * CAP);
  uint32_t* smask = malloc(sizeof(uint32_t) * CAP);
  uint32_t pn[CAP], pm[CAP], ln_[CAP], lm[CAP];
  for (uint64_t g = 0; g < n_groups; ++g) {
    const uint64_t base = g_gbase ? g_gbase[first_group + g * stride] : (first_group + g * stride) * 32;
    const uint64_t gend = g_gbase ? base + g_gcount[first_group + g * stride] : n;
    if (base >= n) break;
    double px[32], py[32], pz[32], lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    uint32_t m0 = 0;
    for (int l = 0; l < 32; ++l) {
      if (SLOT(first_group + g * stride, l) == ~0ull) continue;
      const okd_particle* p = &parts[idx[SLOT(first_group + g * stride, l)]];
      px[l] = p->p[0], py[l] = p->p[1], pz[l] = p->p[2];
      m0 |= 1u << l;
      for (int k = 0; k < 3; ++k) {
        lo[k] = fmin(lo[k], p->p[k]);
        hi[k] = fmax(hi[k], p->p[k]);
      }
    }
    int sp = 0, pend = 0, pleaf = 0, list = 0;
    snode[sp] = 0, smask[sp] = m0, sp++;
    for (;;) {
      const int room = hard - sp - 2 * pend;
      if (defer && (pend >= 32 || (pend > 0 && (room <= 0 || sp == 0)))) {
        const int r = pend < 32 ? pend : 32;
        out->mixed_rounds++;
        out->mixed_round_nodes += r;
        int adds = 0;
        uint32_t ams[32];
        for (int i = 0; i < r; ++i) {
          const uint32_t nd = pn[pend - 1 - i], mk = pm[pend - 1 - i];
          const okd_node* q = &nodes[nd];
          uint32_t am = 0;
          for (int l = 0; l < 32; ++l)
            if ((mk >> l) & 1) {
              const double dx = px[l] - q->u.in.cm[0], dy = py[l] - q->u.in.cm[1], dz = pz[l] - q->u.in.cm[2];
              const double d2 = dx * dx + dy * dy + dz * dz;
              if (q->u.in.size * q->u.in.size < theta2 * d2) am |= 1u << l;
            }
          out->tests += popc(mk);
          ams[i] = am;
          if (am) {
            adds++;
            out->entries++;
            out->lanes += popc(am);
            tl_emit(out, am);
          }
          const uint32_t ms = mk & ~am;
          if (ms) {
            snode[sp] = (uint32_t)q->u.in.right, smask[sp] = ms, sp++;
            snode[sp] = (uint32_t)q->u.in.left, smask[sp] = ms, sp++;
          }
        }
        packstats(out, ams, r, 1);
        pend -= r;
        if (list + adds > 64) out->drains++, list = 0;
        list += adds;
        if (sp > out->max_sp) out->max_sp = sp;
        continue;
      }
      if (defer && (pleaf >= 32 || (pleaf > 0 && sp == 0 && pend == 0))) {
        const int r = pleaf < 32 ? pleaf : 32;
        out->leaf_rounds++;
        out->leaf_round_nodes += r;
        for (int k = 0; k < 8; ++k) {
          int adds = 0;
          uint32_t lms[32];
          for (int i = 0; i < r; ++i) {
            lms[i] = 0;
            const uint32_t nd = ln_[pleaf - 1 - i], mk = lm[pleaf - 1 - i];
            const okd_node* q = &nodes[nd];
            if ((uint64_t)k < q->u.leaf.num_parts) {
              // owner lane removed
              uint32_t m = mk;
              const uint64_t pid = q->u.leaf.leaf_parts[k];
              for (int l = 0; l < 32; ++l)
                if (SLOT(first_group + g * stride, l) != ~0ull && idx[SLOT(first_group + g * stride, l)] == pid) m &= ~(1u << l);
              adds++;
              out->entries++;
              out->part_entries++;
              out->lanes += popc(m);
              lms[i] = m;
              tl_emit(out, m);
            }
          }
          packstats(out, lms, r, 1);
          if (list + adds > 64) out->drains++, list = 0;
          list += adds;
        }
        pleaf -= r;
        continue;
      }
      if (sp == 0) break;
      int nb = sp < 32 ? sp : 32;
      const int rm = room > 1 ? room : 1;
      if (nb > rm) nb = rm;
      sp -= nb;
      out->batches++;
      out->popped += nb;
      int adds = 0;
      uint32_t newn[64], newm[64], fms[32];
      int nn = 0;
      for (int i = 0; i < nb; ++i) {
        const uint32_t nd = snode[sp + i], mk = smask[sp + i];
        const okd_node* q = &nodes[nd];
        fms[i] = 0;
        if (!q->is_internal) {
          out->leaves++;
          if (defer) {
            ln_[pleaf] = nd, lm[pleaf] = mk, pleaf++;
          } else {
            for (uint64_t k = 0; k < q->u.leaf.num_parts; ++k) {
              uint32_t m = mk;
              const uint64_t pid = q->u.leaf.leaf_parts[k];
              for (int l = 0; l < 32; ++l)
                if (SLOT(first_group + g * stride, l) != ~0ull && idx[SLOT(first_group + g * stride, l)] == pid) m &= ~(1u << l);
              out->entries++, out->part_entries++;
              out->lanes += popc(m);
              tl_emit(out, m);
            }
            if (list + (int)q->u.leaf.num_parts > 64) out->drains++, list = 0;
            list += (int)q->u.leaf.num_parts;
          }
          continue;
        }
        double dmin2 = 0, dmax2 = 0;
        for (int k = 0; k < 3; ++k) {
          const double below = lo[k] - q->u.in.cm[k], above = q->u.in.cm[k] - hi[k];
          const double dn = fmax(0.0, fmax(below, above)), df = fmax(fabs(below), fabs(above));
          dmin2 += dn * dn, dmax2 += df * df;
        }
        const double size2 = q->u.in.size * q->u.in.size;
        if (size2 < theta2 * dmin2 * (1 - 1e-9)) {
          out->farn++;
          fms[i] = mk;
          adds++;
          out->entries++;
          out->lanes += popc(mk);
          tl_emit(out, mk);
        } else if (size2 >= theta2 * dmax2 * (1 + 1e-9)) {
          out->nearn++;
          newn[nn] = (uint32_t)q->u.in.right, newm[nn] = mk, nn++;
          newn[nn] = (uint32_t)q->u.in.left, newm[nn] = mk, nn++;
        } else {
          out->mixed++;
          if (defer) {
            pn[pend] = nd, pm[pend] = mk, pend++;
          } else {
            uint32_t am = 0;
            for (int l = 0; l < 32; ++l)
              if ((mk >> l) & 1) {
                const double dx = px[l] - q->u.in.cm[0], dy = py[l] - q->u.in.cm[1], dz = pz[l] - q->u.in.cm[2];
                const double d2 = dx * dx + dy * dy + dz * dz;
                if (size2 < theta2 * d2) am |= 1u << l;
              }
            if (am) {
              out->entries++;
              out->lanes += popc(am);
              tl_emit(out, am);
              if (list + 1 > 64) out->drains++, list = 0;
              list++;
            }
            const uint32_t ms = mk & ~am;
            if (ms) {
              newn[nn] = (uint32_t)q->u.in.right, newm[nn] = ms, nn++;
              newn[nn] = (uint32_t)q->u.in.left, newm[nn] = ms, nn++;
            }
          }
        }
      }
      if (defer) packstats(out, fms, nb, 1);
      if (list + adds > 64) out->drains++, list = 0;
      list += adds;
      for (int i = 0; i < nn; ++i) snode[sp] = newn[i], smask[sp] = newm[i], sp++;
      if (sp > out->max_sp) out->max_sp = sp;
      if (pend > out->max_pend) out->max_pend = pend;
    }
    tl_end_group(out);
    sw_end_group(out);
    ti_end_group();
  }
  free(snode);
  free(smask);
}
