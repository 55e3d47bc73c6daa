// Multi-GPU host side (included by context.cu): symmetric peer buffer, partitioned set-up, distributed
// V-cycle / PCG.  Decomposition: level 0 and the system matrix are row-partitioned by the
// ownership arrays the host hands over (z-slabs of the base lattice); the small patch levels (>= 1) are
// replicated on every GPU so they need no communication; what crosses NVLink is (i) halo rows of the
// PCG / coarse-CG direction vectors, (ii) W x W scalar all-reduces, (iii) two small all-gathers per V-cycle
// (patch defect entries, level-0 values under the patch).

namespace gmg {

static DistPeers peers_of(const gmg_context *h) {
  DistPeers P;
  P.rank = h->dist.rank;
  P.world = h->dist.world;
  for (int q = 0; q < DIST_MAX_RANKS; ++q) P.peer[q] = h->dist.peer[q];
  return P;
}

static int sym_alloc(gmg_context *h, size_t bytes, size_t &off) {
  DistData &d = h->dist;
  off = (d.bump + 255) & ~(size_t)255;
  if (off + bytes > d.bytes) return fail(h, GMG_EINVAL, "symmetric communication buffer too small (gmg_dist_init comm_bytes)");
  d.bump = off + bytes;
  return GMG_OK;
}

template <class T>
static int to_device(gmg_context *h, T *&dst, const std::vector<T> &src) {
  dfree(dst);
  GMG_CUDA(h, dalloc(&dst, (int64_t)src.size()));
  if (!src.empty()) GMG_CUDA(h, copy_sync(h, dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice));
  return GMG_OK;
}

static int dist_check_error(gmg_context *h) {
  int e = 0;
  GMG_CUDA(h, copy_sync(h, &e, h->dist.d_error, sizeof(int), cudaMemcpyDeviceToHost));
  if (e) return fail(h, GMG_ENCCL, "multi-GPU: a wait on a peer flag timed out (a rank is missing or failed)");
  return GMG_OK;
}

static int dist_allreduce1(gmg_context *h, double *v, int k = 1) {
  DistData &d = h->dist;
  dist_allreduce<<<1, 32, 0, h->stream>>>(peers_of(h), v, k, ++d.seq[CH_RED], d.d_error);
  GMG_LAUNCH_CHECK(h);
  return GMG_OK;
}

// push (src -> peers' region) + wait for every source rank on `channel`
static int dist_exchange(gmg_context *h, int n_send, const int *send_src, const unsigned char *send_peer,
                         const int *send_dst, size_t region, const double *src, int channel, uint32_t dst_mask,
                         uint32_t src_mask) {
  DistData &d = h->dist;
  const uint64_t seq = ++d.seq[channel];
  dist_push<<<std::max(cdiv(n_send, 256), 1), 256, 0, h->stream>>>(peers_of(h), n_send, send_src, send_peer, send_dst, region,
                                                                  src, channel, seq, dst_mask, h->counter + 1);
  GMG_LAUNCH_CHECK(h);
  dist_wait<<<1, 32, 0, h->stream>>>(peers_of(h), channel, seq, src_mask, d.d_error);
  GMG_LAUNCH_CHECK(h);
  return GMG_OK;
}

static int dist_halo(gmg_context *h, const DistMat &M, size_t region, int channel) {
  const double *src = reinterpret_cast<const double *>(h->dist.buf + region) + M.n_halo_lo;  // owned part
  return dist_exchange(h, M.n_send, M.send_src, M.send_peer, M.send_dst, region, src, channel, M.dst_mask, M.src_mask);
}

static int dist_gather(gmg_context *h, const GatherPlan &G, const double *src, int channel) {
  const uint32_t all = (1u << h->dist.world) - 1u;
  return dist_exchange(h, G.n_send, G.send_src, G.send_peer, G.send_dst, G.region, src, channel, all, all);
}

static void free_distmat(DistMat &M) {
  dfree(M.rev_src);
  dfree(M.rev_dst);
  dfree(M.rev_peer);
  M.n_rev = 0;
  free_sell(M.A);
  dfree(M.send_src);
  dfree(M.send_dst);
  dfree(M.send_peer);
  dfree(M.send_hpos);
  M = DistMat{};
}
static void free_gather(GatherPlan &G) {
  dfree(G.send_src);
  dfree(G.send_dst);
  dfree(G.send_peer);
  G = GatherPlan{};
}

static int build_distmat(gmg_context *h, const HostCsr &g, const std::vector<int32_t> &owner, DistMat &M, LocalMatrix &lm,
                         ExchangePlan &plan, bool with_reverse = false) {
  DistData &d = h->dist;
  partition_matrix(d.rank, d.world, g.n_rows, g.n_cols, g.rowptr.data(), g.col.data(), g.val.data(), owner.data(),
                   owner.data(), lm, plan);
  HostCsr loc;
  loc.n_rows = lm.n_owned;
  loc.n_cols = lm.n_owned + lm.n_halo;
  loc.rowptr = lm.rowptr;
  loc.col = lm.col;
  loc.val = lm.val;
  free_distmat(M);
  if (int rc = build_sell_host(h, loc, h->drop_tol, M.A)) return rc;
  M.n_owned = lm.n_owned;
  M.n_halo = lm.n_halo;
  M.n_halo_lo = lm.n_halo_lo;
  std::vector<int> ssrc, sdst;
  std::vector<unsigned char> speer;
  for (int q = 0; q < d.world; ++q) {
    if (q == d.rank) continue;
    if (!plan.send_idx[q].empty()) M.dst_mask |= 1u << q;
    if (plan.recv_count[q] > 0) M.src_mask |= 1u << q;
    for (size_t k = 0; k < plan.send_idx[q].size(); ++k) {
      ssrc.push_back(plan.send_idx[q][k]);
      sdst.push_back(plan.send_dst_base[q] + (int)k);
      speer.push_back((unsigned char)q);
    }
  }
  {  // sort by source row: the persistent CG lets the block that owns a row push it
    std::vector<int> order(ssrc.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ssrc[a] < ssrc[b]; });
    std::vector<int> s2(ssrc.size()), d2(ssrc.size());
    std::vector<unsigned char> p2(ssrc.size());
    for (size_t i = 0; i < order.size(); ++i) {
      s2[i] = ssrc[order[i]];
      d2[i] = sdst[order[i]];
      p2[i] = speer[order[i]];
    }
    ssrc.swap(s2);
    sdst.swap(d2);
    speer.swap(p2);
  }
  {
    std::vector<int> hpos(ssrc.size());
    for (size_t i = 0; i < hpos.size(); ++i)
      hpos[i] = sdst[i] - (d.rank > (int)speer[i] ? plan.n_owned_of[speer[i]] : 0);
    if (int rc = to_device(h, M.send_hpos, hpos)) return rc;
  }
  if (with_reverse) {
    int stride = 1;
    for (int q = 0; q < d.world; ++q) stride = std::max(stride, plan.n_owned_of[q] + plan.n_halo_of[q]);
    M.rev_stride = stride;
    // (symmetric allocation: every rank reaches this point with the same sizes)
    if (int rc = sym_alloc(h, sizeof(double) * (size_t)stride * d.world, M.rev_region)) return rc;
    std::vector<int> rsrc, rdst;
    std::vector<unsigned char> rpeer;
    for (int e = 0; e < lm.n_halo; ++e) {
      const int pos = e < lm.n_halo_lo ? e : lm.n_owned + e;  // position in my extended vector
      rsrc.push_back(pos);
      rdst.push_back(d.rank * stride + pos);
      rpeer.push_back((unsigned char)lm.halo_owner[e]);
    }
    M.n_rev = (int)rsrc.size();
    if (int rc = to_device(h, M.rev_src, rsrc)) return rc;
    if (int rc = to_device(h, M.rev_dst, rdst)) return rc;
    if (int rc = to_device(h, M.rev_peer, rpeer)) return rc;
  }
  M.h_send_src = ssrc;
  M.n_send = (int)ssrc.size();
  if (int rc = to_device(h, M.send_src, ssrc)) return rc;
  if (int rc = to_device(h, M.send_dst, sdst)) return rc;
  return to_device(h, M.send_peer, speer);
}

// owned entries (local src index, destination position) pushed to every rank, self included
static int build_gather(gmg_context *h, const std::vector<int> &src_local, const std::vector<int> &dst_pos, int n_total,
                        GatherPlan &G) {
  DistData &d = h->dist;
  free_gather(G);
  std::vector<int> ssrc, sdst;
  std::vector<unsigned char> speer;
  for (int q = 0; q < d.world; ++q)
    for (size_t k = 0; k < src_local.size(); ++k) {
      ssrc.push_back(src_local[k]);
      sdst.push_back(dst_pos[k]);
      speer.push_back((unsigned char)q);
    }
  G.n_total = n_total;
  G.n_send = (int)ssrc.size();
  if (int rc = sym_alloc(h, sizeof(double) * (size_t)std::max(n_total, 1), G.region)) return rc;
  if (int rc = to_device(h, G.send_src, ssrc)) return rc;
  if (int rc = to_device(h, G.send_dst, sdst)) return rc;
  return to_device(h, G.send_peer, speer);
}

static int dist_setup(gmg_context *h) {
  DistData &d = h->dist;
  if (d.hS.empty() || d.hA0.empty() || d.sys_owner.empty() || d.l0_owner.empty())
    return fail(h, GMG_EINVAL, "multi-GPU: system / level-0 matrices and their ownership must be set before gmg_setup");
  if ((int)d.sys_owner.size() != d.hS.n_rows || (int)d.l0_owner.size() != d.hA0.n_rows)
    return fail(h, GMG_EINVAL, "multi-GPU: ownership array size mismatch");
  d.bump = DIST_HEADER_BYTES;
  d.n_sys = d.hS.n_rows;
  d.n_l0 = d.hA0.n_rows;
  h->n_sys = d.n_sys;
  LocalMatrix lmS, lmA;
  ExchangePlan plS, plA;
  int rc;
  {
    TraceScope tr("  dist: system matrix");
    if ((rc = build_distmat(h, d.hS, d.sys_owner, d.S, lmS, plS, true))) return rc;
  }
  {
    TraceScope tr("  dist: level-0 matrix");
    if ((rc = build_distmat(h, d.hA0, d.l0_owner, d.A0, lmA, plA))) return rc;
  }
  d.n_sys_owned = lmS.n_owned;
  d.n_l0_owned = lmA.n_owned;
  // row-pattern format of the rank-local block (columns in [-n_halo_lo, n_owned + halo)), else CSELL, else SELL
  if (h->compress >= 2)
    if ((rc = build_pat(h, d.A0.A))) return rc;
  if (h->compress >= 1 && !d.A0.A.patterned)
    if ((rc = build_csell(h, d.A0.A))) return rc;  // note: offsets into [owned | halo] must fit 16 bits, else plain SELL
  {
    int per_sm = 0;
    const bool pat = d.A0.A.patterned && h->compress >= 2;
    const bool comp = !pat && d.A0.A.compressed && h->compress;
    cudaError_t e = pat    ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_persistent_dist<1024, PatView>, 1024, 0)
                    : comp ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_persistent_dist<1024, CsellView>, 1024, 0)
                           : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_persistent_dist<1024, SellView>, 1024, 0);
    if (e != cudaSuccess || per_sm < 1) return fail(h, GMG_ECUDA, "occupancy query of the distributed CG kernel failed");
    d.cg_grid = h->sm_count * per_sm;
  }
  {  // send-list ranges per block of the persistent CG (same slice partition as the kernel)
    const int nb = d.cg_grid, n_slices = d.A0.A.v.n_slices;
    std::vector<int> bp(nb + 1, 0);
    size_t t = 0;
    for (int b = 0; b < nb; ++b) {
      const int row_end = (int)(((int64_t)n_slices * (b + 1)) / nb) * 32;
      while (t < d.A0.h_send_src.size() && d.A0.h_send_src[t] < row_end) ++t;
      bp[b + 1] = (int)t;
    }
    bp[nb] = (int)d.A0.h_send_src.size();
    if ((rc = to_device(h, d.cg_send_block_ptr, bp))) return rc;
  }
  // symmetric extended vectors: size = max over ranks of owned + halo (every rank knows every size)
  int ext_sys = 0, ext_l0 = 0;
  for (int q = 0; q < d.world; ++q) {
    ext_sys = std::max(ext_sys, plS.n_owned_of[q] + plS.n_halo_of[q]);
    ext_l0 = std::max(ext_l0, plA.n_owned_of[q] + plA.n_halo_of[q]);
  }
  if ((rc = sym_alloc(h, sizeof(double) * (size_t)ext_sys, d.reg_pcg_x))) return rc;
  if ((rc = sym_alloc(h, sizeof(double) * (size_t)ext_sys, d.reg_pcg_d))) return rc;
  if ((rc = sym_alloc(h, sizeof(double) * (size_t)ext_l0, d.reg_cg_d))) return rc;
  {
    int halo_max = 0;
    for (int q = 0; q < d.world; ++q) halo_max = std::max(halo_max, plA.n_halo_of[q]);
    if ((rc = sym_alloc(h, 16 * (size_t)std::max(halo_max, 1), d.reg_cg_ll))) return rc;
    if ((rc = sym_alloc(h, DIST_CG_SLOT_BYTES, d.reg_cg_slots))) return rc;
    if (d.cg_grid > DIST_CG_MAXB) return fail(h, GMG_EINVAL, "distributed CG: more blocks than reduction slots");
  }
  if ((rc = to_device(h, d.sys_owned_global, lmS.owned_global))) return rc;
  std::vector<int> sys_g2l(d.n_sys, -1), l0_g2l(d.n_l0, -1);
  for (int i = 0; i < lmS.n_owned; ++i) sys_g2l[lmS.owned_global[i]] = i;
  for (int i = 0; i < lmA.n_owned; ++i) l0_g2l[lmA.owned_global[i]] = i;

  // ---- copy indices
  const int nl = h->n_levels;
  if ((int)d.h_copy_g.size() != nl) return fail(h, GMG_EINVAL, "multi-GPU: copy indices missing");
  {
    std::vector<int> cs, cl;
    for (size_t k = 0; k < d.h_copy_g[0].size(); ++k) {
      const int g = d.h_copy_g[0][k], lv = d.h_copy_l[0][k];
      if (d.sys_owner[g] != d.rank) continue;
      if (d.l0_owner[lv] != d.rank) return fail(h, GMG_EINVAL, "multi-GPU: a system dof and its level-0 twin have different owners");
      cs.push_back(sys_g2l[g]);
      cl.push_back(l0_g2l[lv]);
    }
    d.n_copy0 = (int)cs.size();
    if ((rc = to_device(h, d.copy0_sys, cs))) return rc;
    if ((rc = to_device(h, d.copy0_l0, cl))) return rc;
  }
  {
    std::vector<int> src, pos;
    d.gather_g_offset.assign(nl + 1, 0);
    for (auto p : d.from_sys) dfree(p);
    for (auto p : d.from_lvl) dfree(p);
    d.from_sys.assign(nl, nullptr);
    d.from_lvl.assign(nl, nullptr);
    d.n_from.assign(nl, 0);
    int total = 0;
    for (int l = 1; l < nl; ++l) {
      d.gather_g_offset[l] = total;
      std::vector<int> fs, fl;
      for (size_t k = 0; k < d.h_copy_g[l].size(); ++k) {
        const int g = d.h_copy_g[l][k];
        if (d.sys_owner[g] == d.rank) {
          src.push_back(sys_g2l[g]);
          pos.push_back(total + (int)k);
          fs.push_back(sys_g2l[g]);
          fl.push_back(d.h_copy_l[l][k]);
        }
      }
      total += (int)d.h_copy_g[l].size();
      d.n_from[l] = (int)fs.size();
      if ((rc = to_device(h, d.from_sys[l], fs))) return rc;
      if ((rc = to_device(h, d.from_lvl[l], fl))) return rc;
    }
    d.gather_g_offset[nl] = total;
    if ((rc = build_gather(h, src, pos, total, d.gather_g))) return rc;
  }
  // ---- level 0 <-> level 1 transfers
  free_sell(d.R0);
  free_sell(d.P0F);
  free_gather(d.gather_c);
  if (nl > 1) {
    const HostCsr &P = h->levels[0].hP;
    if (P.empty()) return fail(h, GMG_EINVAL, "prolongation from level 0 missing");
    // restriction: rows = my level-0 dofs, columns = level-1 dofs (replicated input)
    HostCsr R = transpose(P);
    HostCsr Rl;
    Rl.n_cols = R.n_cols;
    extract_owned_rows(d.rank, R.n_rows, R.rowptr.data(), R.col.data(), R.val.data(), d.l0_owner.data(), Rl.rowptr, Rl.col,
                       Rl.val);
    Rl.n_rows = (int)Rl.rowptr.size() - 1;
    if ((rc = build_sell_host(h, Rl, 0.0, d.R0))) return rc;
    // prolongation through the footprint F = level-0 dofs that are parents of level-1 dofs
    std::vector<int> F(P.col.begin(), P.col.end());
    std::sort(F.begin(), F.end());
    F.erase(std::unique(F.begin(), F.end()), F.end());
    std::vector<int> posF(d.n_l0, -1);
    for (size_t k = 0; k < F.size(); ++k) posF[F[k]] = (int)k;
    HostCsr PF = P;
    PF.n_cols = (int)F.size();
    for (auto &c : PF.col) c = posF[c];
    if ((rc = build_sell_host(h, PF, 0.0, d.P0F))) return rc;
    std::vector<int> src, pos;
    for (size_t k = 0; k < F.size(); ++k)
      if (d.l0_owner[F[k]] == d.rank) {
        src.push_back(l0_g2l[F[k]]);
        pos.push_back((int)k);
      }
    if ((rc = build_gather(h, src, pos, (int)F.size(), d.gather_c))) return rc;
  }
  // ---- solution all-gather (global positions)
  {
    std::vector<int> src(lmS.n_owned), pos(lmS.owned_global.begin(), lmS.owned_global.end());
    for (int i = 0; i < lmS.n_owned; ++i) src[i] = i;
    if ((rc = build_gather(h, src, pos, d.n_sys, d.gather_x))) return rc;
  }
  // ---- work vectors
  for (double **p : {&d.l0_defect, &d.l0_sol, &d.cg_g, &d.cg_h}) {
    dfree(*p);
    GMG_CUDA(h, dalloc(p, d.n_l0_owned));
    GMG_CUDA(h, cudaMemsetAsync(*p, 0, sizeof(double) * std::max(d.n_l0_owned, 1), h->stream));
  }
  for (double **p : {&d.g, &d.hh}) {
    dfree(*p);
    GMG_CUDA(h, dalloc(p, d.n_sys_owned));
  }
  dfree(d.cg_partials);
  GMG_CUDA(h, dalloc(&d.cg_partials, d.cg_grid));
  GMG_CUDA(h, cudaMemsetAsync(d.buf + DIST_HEADER_BYTES, 0, d.bump - DIST_HEADER_BYTES, h->stream));
  GMG_CUDA(h, cudaStreamSynchronize(h->stream));
  return GMG_OK;
}

static int dist_coarse_cg(gmg_context *h, const double *b, double *x) {
  DistData &d = h->dist;
  const int slot = h->cg_cursor % h->cg_ring;
  h->cg_cursor++;
  CgResult *res = h->cg_results + slot;
  SellView v = d.A0.A.v;
  CsellView cv = d.A0.A.cv;
  PatView pv = d.A0.A.pv;
  const bool pat = d.A0.A.patterned && h->compress >= 2;
  const bool comp = !pat && d.A0.A.compressed && h->compress;
  DistCgArgs D;
  D.P = peers_of(h);
  D.n_owned = d.A0.n_owned;
  D.n_halo = d.A0.n_halo;
  D.n_halo_lo = d.A0.n_halo_lo;
  D.n_send = d.A0.n_send;
  D.send_src = d.A0.send_src;
  D.send_peer = d.A0.send_peer;
  D.send_hpos = d.A0.send_hpos;
  D.send_block_ptr = d.cg_send_block_ptr;
  D.region_d = d.reg_cg_d;
  D.region_ll = d.reg_cg_ll;
  D.region_slots = d.reg_cg_slots;
  D.prof = h->cg_prof;
  // 12-bit launch id in the tag: a tag repeats only after 4096 launches, by when every slot it was used on has been
  // overwritten hundreds of times (each launch rewrites its reduction slots and the whole halo area); 0 is skipped in a
  // way that keeps the parity alternating (ids run 1, 2, ..., 4095, 4098 -> 2 would repeat: skip two at the wrap)
  d.launch_id++;
  if ((d.launch_id & 0xfffu) == 0) d.launch_id += 2;
  D.tag_base = (uint32_t)((d.launch_id & 0xfffu) << 20);
  int max_it = h->coarse_max_it;
  double tol = h->coarse_tol;
  int grid = d.cg_grid;
  void *args[] = {pat ? (void *)&pv : comp ? (void *)&cv : (void *)&v, (void *)&b, &x, &d.cg_g, &d.cg_h, &d.cg_partials, &max_it,
                  &tol, &res, &D, &d.d_error};
  int ev = -1;
  if (h->ev_used < (int)h->ev_begin.size()) {
    ev = h->ev_used++;
    cudaEventRecord(h->ev_begin[ev], h->stream);
  }
  if (pat)
    GMG_CUDA(h, cudaLaunchCooperativeKernel((void *)cg_persistent_dist<1024, PatView>, dim3(grid), dim3(1024), args, 0,
                                            h->stream));
  else if (comp)
    GMG_CUDA(h, cudaLaunchCooperativeKernel((void *)cg_persistent_dist<1024, CsellView>, dim3(grid), dim3(1024), args, 0,
                                            h->stream));
  else
    GMG_CUDA(h, cudaLaunchCooperativeKernel((void *)cg_persistent_dist<1024, SellView>, dim3(grid), dim3(1024), args, 0,
                                            h->stream));
  h->launches++;
  if (ev >= 0) {
    cudaEventRecord(h->ev_end[ev], h->stream);
    h->ev_result_slot[ev] = slot;
  }
  return GMG_OK;
}

// PreconditionMG::vmult, distributed: src / dst are owned-only system vectors.  Split into the parts that talk to
// peers (sequence-numbered, launched directly) and the purely local sweeps (replayed as CUDA graphs).
static int dist_vc_pre(gmg_context *h, const double *src) {  // copy_to_mg
  DistData &d = h->dist;
  const int nl = h->n_levels;
  GMG_CUDA(h, cudaMemsetAsync(d.l0_defect, 0, sizeof(double) * std::max(d.n_l0_owned, 1), h->stream));
  if (d.n_copy0) {
    vec_gather<<<cdiv(d.n_copy0, 256), 256, 0, h->stream>>>(d.n_copy0, d.copy0_l0, d.copy0_sys, src, d.l0_defect);
    GMG_LAUNCH_CHECK(h);
  }
  if (nl > 1)
    if (int rc = dist_gather(h, d.gather_g, src, CH_GATHER_G)) return rc;
  return GMG_OK;
}

static int dist_vc_down(gmg_context *h) {
  DistData &d = h->dist;
  const int nl = h->n_levels;
  const double *gg = reinterpret_cast<const double *>(d.buf + d.gather_g.region);
  for (int l = 1; l < nl; ++l) {
    Level &L = h->levels[l];
    GMG_CUDA(h, cudaMemsetAsync(L.defect, 0, sizeof(double) * std::max(L.n, 1), h->stream));
    if (L.n_copy) {
      vec_gather_from<<<cdiv(L.n_copy, 256), 256, 0, h->stream>>>(L.n_copy, L.copy_l, gg + d.gather_g_offset[l], L.defect);
      GMG_LAUNCH_CHECK(h);
    }
  }
  for (int l = nl - 1; l >= 1; --l) {
    Level &L = h->levels[l];
    if (int rc = smooth(h, L, L.sol, L.defect, true)) return rc;
    if (int rc = spmv<EPI_RESID, DOT_NONE>(h, L.AI, L.sol, L.t, L.defect)) return rc;
    if (l > 1) {
      if (int rc = spmv<EPI_ADD, DOT_NONE>(h, h->levels[l - 1].R, L.t, h->levels[l - 1].defect)) return rc;
    } else {
      if (int rc = spmv<EPI_ADD, DOT_NONE>(h, d.R0, L.t, d.l0_defect)) return rc;
    }
  }
  return GMG_OK;
}

static int dist_vc_up(gmg_context *h, double *dst) {
  DistData &d = h->dist;
  const int nl = h->n_levels;
  for (int l = 1; l < nl; ++l) {
    Level &L = h->levels[l];
    if (l > 1) {
      if (int rc = spmv<EPI_ADD, DOT_NONE>(h, h->levels[l - 1].P, h->levels[l - 1].sol, L.sol)) return rc;
    } else {
      const double *c = reinterpret_cast<const double *>(d.buf + d.gather_c.region);
      if (int rc = spmv<EPI_ADD, DOT_NONE>(h, d.P0F, c, L.sol)) return rc;
    }
    if (L.IT.valid && L.IT.stored_nnz > 0)
      if (int rc = spmv<EPI_SUB, DOT_NONE>(h, L.IT, L.sol, L.defect)) return rc;
    if (int rc = smooth(h, L, L.sol, L.defect, false)) return rc;
  }
  // copy_from_mg
  GMG_CUDA(h, cudaMemsetAsync(dst, 0, sizeof(double) * std::max(d.n_sys_owned, 1), h->stream));
  if (d.n_copy0) {
    vec_gather<<<cdiv(d.n_copy0, 256), 256, 0, h->stream>>>(d.n_copy0, d.copy0_sys, d.copy0_l0, d.l0_sol, dst);
    GMG_LAUNCH_CHECK(h);
  }
  for (int l = 1; l < nl; ++l)
    if (d.n_from[l]) {
      vec_gather<<<cdiv(d.n_from[l], 256), 256, 0, h->stream>>>(d.n_from[l], d.from_sys[l], d.from_lvl[l], h->levels[l].sol, dst);
      GMG_LAUNCH_CHECK(h);
    }
  return GMG_OK;
}

static int dist_vcycle(gmg_context *h, const double *src, double *dst) {
  DistData &d = h->dist;
  const int nl = h->n_levels;
  const bool graphs = h->use_graphs && nl > 1 && h->smoother != GMG_SMOOTHER_JACOBI;
  gmg_context::VcGraph *g = nullptr;
  if (graphs) {
    for (auto &c : h->vc_graphs)
      if (c.src == src && c.dst == dst) g = &c;
    if (!g && h->vc_graphs.size() < 8) {
      gmg_context::VcGraph ng;
      ng.src = src;
      ng.dst = dst;
      if (capture_graph(h, [&]() { return dist_vc_down(h); }, ng.down, ng.n_down) == GMG_OK &&
          capture_graph(h, [&]() { return dist_vc_up(h, dst); }, ng.up, ng.n_up) == GMG_OK) {
        h->vc_graphs.push_back(ng);
        g = &h->vc_graphs.back();
      } else {
        if (ng.down) cudaGraphExecDestroy(ng.down);
        if (ng.up) cudaGraphExecDestroy(ng.up);
        h->use_graphs = false;
      }
    }
  }
  if (int rc = dist_vc_pre(h, src)) return rc;
  if (g) {
    GMG_CUDA(h, cudaGraphLaunch(g->down, h->stream));
    h->launches += g->n_down;
  } else if (int rc = dist_vc_down(h)) {
    return rc;
  }
  if (int rc = dist_coarse_cg(h, d.l0_defect, d.l0_sol)) return rc;
  if (nl > 1)
    if (int rc = dist_gather(h, d.gather_c, d.l0_sol, CH_GATHER_C)) return rc;
  if (g) {
    GMG_CUDA(h, cudaGraphLaunch(g->up, h->stream));
    h->launches += g->n_up;
    return GMG_OK;
  }
  return dist_vc_up(h, dst);
}

// distributed SolverCG with the GMG preconditioner; b, x are GLOBAL-length device vectors (every rank holds
// the same b and receives the full solution)
static int dist_pcg(gmg_context *h, const double *b_global, double *x_global, int max_it, double tol, int *iters,
                    double *res0_out, double *res_out) {
  DistData &d = h->dist;
  const int n = d.n_sys_owned;
  h->cg_solve_begin = h->cg_cursor;
  // extended vectors [lower halo | owned | upper halo]; kernels address them from the first owned entry
  double *x = reinterpret_cast<double *>(d.buf + d.reg_pcg_x) + d.S.n_halo_lo;
  double *dd = reinterpret_cast<double *>(d.buf + d.reg_pcg_d) + d.S.n_halo_lo;
  double *bl = h->hh;  // owned part of b (h->hh has global length >= owned)
  PcgScalars hs;
  const int tg = cdiv(std::max(n, 1), 256);
  vec_take<<<tg, 256, 0, h->stream>>>(n, d.sys_owned_global, b_global, bl);
  GMG_LAUNCH_CHECK(h);
  vec_take<<<tg, 256, 0, h->stream>>>(n, d.sys_owned_global, x_global, x);
  GMG_LAUNCH_CHECK(h);
  const int rg = reduce_grid(h, n);
  auto sync_scalars = [&]() -> int {
    GMG_CUDA(h, copy(h, &hs, h->scalars, sizeof(hs), cudaMemcpyDeviceToHost));
    GMG_CUDA(h, cudaStreamSynchronize(h->stream));
    return dist_check_error(h);
  };
  auto finish = [&]() -> int {
    if (int rc = dist_gather(h, d.gather_x, x, CH_GATHER_X)) return rc;
    GMG_CUDA(h, copy(h, x_global, d.buf + d.gather_x.region, sizeof(double) * d.n_sys, cudaMemcpyDeviceToDevice));
    GMG_CUDA(h, cudaStreamSynchronize(h->stream));
    return dist_check_error(h);
  };
  if (int rc = dist_halo(h, d.S, d.reg_pcg_x, CH_HALO_SYS_X)) return rc;
  if (int rc = spmv<EPI_NRESID, DOT_YY>(h, d.S.A, x, d.g, bl, nullptr, 0.0, &h->scalars->res2)) return rc;
  if (int rc = dist_allreduce1(h, &h->scalars->res2)) return rc;
  if (int rc = sync_scalars()) return rc;
  double res = std::sqrt(hs.res2);
  *res0_out = res;
  *res_out = res;
  *iters = 0;
  if (res <= tol) return finish();
  int slot = 0;
  if (int rc = dist_vcycle(h, d.g, d.hh)) return rc;
  pcg_init_direction<<<rg, 256, 0, h->stream>>>(n, d.g, d.hh, dd, h->scalars, slot, h->partials, h->counter);
  GMG_LAUNCH_CHECK(h);
  if (int rc = dist_allreduce1(h, &h->scalars->gh[slot])) return rc;
  int it = 0;
  while (true) {
    ++it;
    if (int rc = dist_halo(h, d.S, d.reg_pcg_d, CH_HALO_SYS_D)) return rc;
    if (int rc = spmv<EPI_ASSIGN, DOT_XY>(h, d.S.A, dd, d.hh, nullptr, nullptr, 0.0, &h->scalars->dh)) return rc;
    if (int rc = dist_allreduce1(h, &h->scalars->dh)) return rc;
    pcg_update<<<rg, 256, 0, h->stream>>>(n, x, d.g, dd, d.hh, h->scalars, slot, h->partials, h->counter);
    GMG_LAUNCH_CHECK(h);
    if (int rc = dist_allreduce1(h, &h->scalars->res2)) return rc;
    if (int rc = sync_scalars()) return rc;
    {
      CgResult r;
      GMG_CUDA(h, copy_sync(h, &r, h->cg_results + ((h->cg_cursor - 1) % h->cg_ring), sizeof(r), cudaMemcpyDeviceToHost));
      if (r.status == 1)
        return fail(h, GMG_ENOCONVERGENCE, "coarse-grid CG: convergence failure in step " + std::to_string(r.iterations));
      if (r.status == 2) return fail(h, GMG_ENCCL, "multi-GPU coarse CG aborted: a peer did not arrive");
    }
    res = std::sqrt(hs.res2);
    *res_out = res;
    *iters = it;
    if (res <= tol) break;
    if (it >= max_it || std::isnan(res))
      return fail(h, GMG_ENOCONVERGENCE, "Iterative method reported convergence failure in step " + std::to_string(it));
    if (int rc = dist_vcycle(h, d.g, d.hh)) return rc;
    vec_dot<<<rg, 256, 0, h->stream>>>(n, d.g, d.hh, &h->scalars->gh[slot ^ 1], h->partials, h->counter);
    GMG_LAUNCH_CHECK(h);
    if (int rc = dist_allreduce1(h, &h->scalars->gh[slot ^ 1])) return rc;
    slot ^= 1;
    pcg_new_direction<<<cdiv(std::max(n, 1), 256), 256, 0, h->stream>>>(n, dd, d.hh, h->scalars, slot);
    GMG_LAUNCH_CHECK(h);
  }
  return finish();
}

// L1 / Linf / Frobenius norm of a row-partitioned matrix: row sums and squares are local; the column sums a rank
// accumulates for its halo columns travel back to the columns' owners (reverse halo exchange), maxima and the sum of
// squares are combined over the ranks in rank order (the same bits on every rank).
static int dist_matrix_norms(gmg_context *h, DistMat &M, double out[3]) {
  DistData &d = h->dist;
  if (!M.A.valid || M.rev_stride == 0) return fail(h, GMG_EINVAL, "matrix norms: no reverse exchange plan for this matrix");
  const int n = M.n_owned, ext = M.n_owned + M.n_halo;
  const int grid = cdiv(std::max(n, 1), 256);
  if (grid > h->partials_cap) return fail(h, GMG_EINVAL, "matrix too large for the partials buffer");
  if (int rc = ensure_stage(h, ext)) return rc;
  double *colsum = h->stage_a;
  GMG_CUDA(h, cudaMemsetAsync(colsum, 0, sizeof(double) * std::max(ext, 1), h->stream));
  sell_norm_partials<<<grid, 256, 0, h->stream>>>(M.A.v, colsum + M.n_halo_lo, h->partials, h->partials + h->partials_cap);
  GMG_LAUNCH_CHECK(h);
  if (int rc = dist_exchange(h, M.n_rev, M.rev_src, M.rev_peer, M.rev_dst, M.rev_region, colsum, CH_REV, M.src_mask, M.dst_mask))
    return rc;
  if (M.n_send > 0) {
    dist_rev_add<<<cdiv(M.n_send, 256), 256, 0, h->stream>>>(M.n_send, M.send_src, M.send_peer, M.send_dst,
                                                             reinterpret_cast<const double *>(d.buf + M.rev_region), M.rev_stride,
                                                             colsum + M.n_halo_lo);
    GMG_LAUNCH_CHECK(h);
  }
  const int g2 = std::min(cdiv(std::max(n, 1), 256), h->partials_cap);
  vec_max_partials<<<g2, 256, 0, h->stream>>>(n, colsum + M.n_halo_lo, h->partials + 2 * h->partials_cap);
  GMG_LAUNCH_CHECK(h);
  std::vector<double> rowmax(grid), frob(grid), colmax(g2);
  GMG_CUDA(h, copy(h, rowmax.data(), h->partials, sizeof(double) * grid, cudaMemcpyDeviceToHost));
  GMG_CUDA(h, copy(h, frob.data(), h->partials + h->partials_cap, sizeof(double) * grid, cudaMemcpyDeviceToHost));
  GMG_CUDA(h, copy_sync(h, colmax.data(), h->partials + 2 * h->partials_cap, sizeof(double) * g2, cudaMemcpyDeviceToHost));
  double mine[3] = {n > 0 ? *std::max_element(colmax.begin(), colmax.end()) : 0.0,
                    n > 0 ? *std::max_element(rowmax.begin(), rowmax.end()) : 0.0, 0.0};
  for (double v : frob) mine[2] += v;
  // every rank's three numbers on every rank: sums of vectors that are zero outside the contributing rank's slot
  std::vector<double> all(3 * (size_t)d.world, 0.0);
  double *dv = nullptr;
  GMG_CUDA(h, dalloc(&dv, 4));
  for (int r = 0; r < d.world; ++r) {
    double v[4] = {0, 0, 0, 0};
    if (r == d.rank) std::copy(mine, mine + 3, v);
    GMG_CUDA(h, copy(h, dv, v, sizeof(v), cudaMemcpyHostToDevice));
    if (int rc = dist_allreduce1(h, dv, 3)) return rc;
    GMG_CUDA(h, copy_sync(h, &all[3 * (size_t)r], dv, sizeof(double) * 3, cudaMemcpyDeviceToHost));
  }
  dfree(dv);
  if (int rc = dist_check_error(h)) return rc;
  out[0] = out[1] = out[2] = 0.0;
  for (int r = 0; r < d.world; ++r) {
    out[0] = std::max(out[0], all[3 * r]);
    out[1] = std::max(out[1], all[3 * r + 1]);
    out[2] += all[3 * r + 2];
  }
  out[2] = std::sqrt(out[2]);
  return GMG_OK;
}

static void dist_free(gmg_context *h) {
  DistData &d = h->dist;
  free_distmat(d.S);
  free_distmat(d.A0);
  free_gather(d.gather_g);
  free_gather(d.gather_c);
  free_gather(d.gather_x);
  free_sell(d.R0);
  free_sell(d.P0F);
  dfree(d.sys_owned_global);
  dfree(d.copy0_sys);
  dfree(d.copy0_l0);
  for (auto p : d.from_sys) dfree(p);
  for (auto p : d.from_lvl) dfree(p);
  dfree(d.l0_defect);
  dfree(d.l0_sol);
  dfree(d.cg_g);
  dfree(d.cg_h);
  dfree(d.g);
  dfree(d.hh);
  dfree(d.cg_partials);
  dfree(d.cg_send_block_ptr);
  dfree(d.d_error);
  for (int q = 0; q < d.world; ++q)
    if (q != d.rank && d.peer[q]) cudaIpcCloseMemHandle(d.peer[q]);
  if (d.buf) cudaFree(d.buf);
  d = DistData{};
}

}  // namespace gmg

extern "C" {

int gmg_dist_init(gmg_handle h, int rank, int world, int64_t comm_bytes, void *ipc_handle_out) {
  if (!h || world < 1 || world > DIST_MAX_RANKS || rank < 0 || rank >= world || !ipc_handle_out) return GMG_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  gmg::enter(h);
  DistData &d = h->dist;
  if (d.buf) return fail(h, GMG_EINVAL, "gmg_dist_init called twice");
  d.rank = rank;
  d.world = world;
  d.bytes = (size_t)std::max<int64_t>(comm_bytes, (int64_t)DIST_HEADER_BYTES * 2);
  GMG_CUDA(h, cudaMalloc((void **)&d.buf, d.bytes));  // plain cudaMalloc: pool memory cannot be IPC-exported
  GMG_CUDA(h, cudaMemset(d.buf, 0, d.bytes));
  GMG_CUDA(h, cudaDeviceSynchronize());
  cudaIpcMemHandle_t hd;
  GMG_CUDA(h, cudaIpcGetMemHandle(&hd, d.buf));
  std::memcpy(ipc_handle_out, &hd, sizeof(hd));
  GMG_CUDA(h, dalloc(&d.d_error, 1));
  GMG_CUDA(h, cudaMemsetAsync(d.d_error, 0, sizeof(int), h->stream));
  GMG_CUDA(h, cudaStreamSynchronize(h->stream));
  d.bump = DIST_HEADER_BYTES;
  return GMG_OK;
}

int gmg_dist_connect(gmg_handle h, const void *all_handles) {
  if (!h || !all_handles || !h->dist.buf) return GMG_EINVAL;
  gmg::enter(h);
  DistData &d = h->dist;
  for (int q = 0; q < d.world; ++q) {
    if (q == d.rank) {
      d.peer[q] = d.buf;
      continue;
    }
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, (const char *)all_handles + 64 * (size_t)q, sizeof(hd));
    void *p = nullptr;
    GMG_CUDA(h, cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    d.peer[q] = (char *)p;
  }
  d.on = d.world > 1;
  return GMG_OK;
}

int gmg_set_ownership(gmg_handle h, int which, int level, int32_t n, const int32_t *owner) {
  if (!h || !owner || n < 0) return GMG_EINVAL;
  if (which == GMG_SYSTEM) h->dist.sys_owner.assign(owner, owner + n);
  else if (which == GMG_LEVEL && level == 0) h->dist.l0_owner.assign(owner, owner + n);
  else return fail(h, GMG_EINVAL, "ownership is defined for the system matrix and level 0 (patch levels are replicated)");
  h->is_setup = false;
  rhs_invalidate_partition(h);
  return GMG_OK;
}

int gmg_dist_rank(gmg_handle h, int *rank, int *world) {
  if (!h) return GMG_EINVAL;
  if (rank) *rank = h->dist.rank;
  if (world) *world = h->dist.world;
  return GMG_OK;
}

// CPU-only probe of the partitioner (tests): local CSR + maps of `rank` in malloc'ed arrays (gmg_free_host)
int gmg_partition_probe(int rank, int world, int32_t n_rows, const int64_t *rowptr, const int32_t *col, const double *val,
                        const int32_t *owner, int32_t *n_owned, int32_t *n_halo, int64_t **l_rowptr, int32_t **l_col,
                        double **l_val, int32_t **owned_global, int32_t **halo_global, int32_t **send_count,
                        int32_t **send_idx, int32_t **send_dst_base) {
  try {
    LocalMatrix lm;
    ExchangePlan pl;
    partition_matrix(rank, world, n_rows, n_rows, rowptr, col, val, owner, owner, lm, pl);
    auto dup = [](const auto &v, auto **out) {
      using T = typename std::remove_reference<decltype(v[0])>::type;
      using U = typename std::remove_const<T>::type;
      *out = (U *)std::malloc(sizeof(U) * std::max<size_t>(v.size(), 1));
      std::copy(v.begin(), v.end(), *out);
    };
    *n_owned = lm.n_owned;
    *n_halo = lm.n_halo;
    dup(lm.rowptr, l_rowptr);
    dup(lm.col, l_col);
    dup(lm.val, l_val);
    dup(lm.owned_global, owned_global);
    dup(lm.halo_global, halo_global);
    std::vector<int32_t> cnt(world + 1), flat;  // cnt[world] = n_halo_lo
    for (int q = 0; q < world; ++q) {
      cnt[q] = (int32_t)pl.send_idx[q].size();
      flat.insert(flat.end(), pl.send_idx[q].begin(), pl.send_idx[q].end());
    }
    cnt[world] = lm.n_halo_lo;
    dup(cnt, send_count);
    dup(flat, send_idx);
    dup(pl.send_dst_base, send_dst_base);
  } catch (std::exception &) {
    return GMG_EINVAL;
  }
  return GMG_OK;
}

void gmg_free_host(void *p) { std::free(p); }

// developer probe: NVLink round-trip latency of tagged LL words between ranks 0 and 1 (both ranks must call)
int gmg_dist_pingpong(gmg_handle h, int iters, int mode, double *us_per_round_trip) {
  if (!h || !h->dist.on || h->dist.world < 2) return GMG_EINVAL;
  gmg::enter(h);
  DistData &d = h->dist;
  if (d.rank > 1) { if (us_per_round_trip) *us_per_round_trip = 0.0; return GMG_OK; }
  static size_t region = 0;
  if (region == 0) { region = DIST_HEADER_BYTES - 64; }
  long long *out = nullptr;
  GMG_CUDA(h, dalloc(&out, 1));
  static uint32_t tb = 0x40000000u;
  tb += 0x100000u;
  dist_pingpong<<<1, 32, 0, h->stream>>>(peers_of(h), region, iters, mode, tb, out);
  GMG_LAUNCH_CHECK(h);
  long long cyc = 0;
  GMG_CUDA(h, copy_sync(h, &cyc, out, sizeof(cyc), cudaMemcpyDeviceToHost));
  dfree(out);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, h->device);
  if (us_per_round_trip) *us_per_round_trip = (double)cyc / (khz * 1e-3) / iters;
  return GMG_OK;
}

}  // extern "C"
