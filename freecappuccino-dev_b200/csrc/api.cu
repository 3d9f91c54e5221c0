// api.cu -- the extern "C" entry points of include/fcp.h: context creation (mesh upload, create_CSR_matrix,
// face gather lists), field transfer, operator dispatch, calcp_simple driver.
#include <algorithm>
#include <cstring>
#include <thread>
#include "fcp_internal.h"

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
template <class Fn> static void parallel_for(int64_t n, Fn fn) {
  unsigned hw = std::thread::hardware_concurrency();
  int nt = (int)std::max(1u, std::min(hw ? hw : 1u, 32u));
  if (n < 200000) nt = 1;
  if (nt == 1) { fn((int64_t)0, n); return; }
  std::vector<std::thread> th;
  int64_t per = (n + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    int64_t b = t * per, e = std::min(n, b + per);
    if (b >= e) break;
    th.emplace_back([=]() { fn(b, e); });
  }
  for (auto &t : th) t.join();
}

static int select_device(int device) {
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    fcp_set_error("no CUDA device visible (%s); libfcp_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
    return FCP_ENODEVICE;
  }
  if (device < 0 || device >= ndev) {
    fcp_set_error("device %d out of range (0..%d)", device, ndev - 1);
    return FCP_EINVAL;
  }
  cudaDeviceProp prop;
  FCP_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    fcp_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return FCP_ENODEVICE;
  }
  FCP_CUDA(cudaSetDevice(device));
  return FCP_OK;
}

static int field_count(const fcp_ctx *c, int field, int64_t *count) {
  if (field < 0 || field >= FCP_F_COUNT) { fcp_set_error("bad field id %d", field); return FCP_EINVAL; }
  if (fcp_is_gradient_field(field)) *count = 3 * (int64_t)c->nT;
  else if (field == FCP_F_FLMASS) *count = c->nF;
  else if (field == FCP_F_A || field == FCP_F_H) *count = c->pat.nnzp;   // device storage is SELL; host-visible count is nnz
  else if (field == FCP_F_APR) *count = c->npro;
  else *count = c->nT;
  return FCP_OK;
}
static int field_ptr(fcp_ctx *c, int field, double **p) {
  int64_t cnt = 0;
  FCP_TRY(field_count(c, field, &cnt));
  if (!c->field[field]) {
    FCP_TRY(dev_alloc(&c->field[field], (size_t)cnt));
    FCP_CUDA(cudaMemsetAsync(c->field[field], 0, sizeof(double) * (size_t)std::max<int64_t>(cnt, 1), c->stream));
  }
  *p = c->field[field];
  return FCP_OK;
}
#define FIELD(var, id) double *var = nullptr; FCP_TRY(field_ptr(ctx, id, &var))

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
// fills a freshly constructed context; on any error the caller (fcp_ctx_create) destroys it, so nothing leaks on the error paths
static int ctx_build(fcp_ctx *c, const fcp_mesh_desc *md, int device) {
  c->device = device;
  c->n = md->numCells; c->F = md->numInnerFaces; c->B = md->numBoundaryFaces; c->nb = md->numBoundaries;
  c->nT = c->n + c->B; c->nF = c->F + c->B;
  const int32_t n = c->n, F = c->F, B = c->B;
  FCP_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  FCP_CUDA(cudaEventCreate(&c->t0));
  FCP_CUDA(cudaEventCreate(&c->t1));
  c->bctype.assign(md->bctype, md->bctype + c->nb);
  c->nfaces.assign(md->nfaces, md->nfaces + c->nb);
  c->startFace.assign(md->startFace, md->startFace + c->nb);

  // per-boundary-face patch type; validate the patch table
  std::vector<int32_t> bft(std::max(B, 1), FCP_BC_WALL);
  for (int32_t ib = 0; ib < c->nb; ++ib) {
    if (c->startFace[ib] < F || c->startFace[ib] + c->nfaces[ib] > c->nF) {
      fcp_set_error("patch %d: faces [%d,%d) outside the boundary range [%d,%d)", ib, c->startFace[ib], c->startFace[ib] + c->nfaces[ib], F, c->nF);
      return FCP_EINVAL;
    }
    for (int32_t i = 0; i < c->nfaces[ib]; ++i) bft[c->startFace[ib] - F + i] = c->bctype[ib];
    if (c->bctype[ib] == FCP_BC_PRESSURE) c->has_pressure_patch = true;
    if (c->bctype[ib] == FCP_BC_OUTLET) c->has_outlet = true;
    if (c->bctype[ib] == FCP_BC_PROCESS) c->npro += c->nfaces[ib];
  }
  c->g_pressure_patch = c->has_pressure_patch;   // fcp_comm_init replaces these by the flags over all ranks
  c->g_outlet = c->has_outlet;
  for (int32_t f = 0; f < c->nF; ++f) {
    if (md->owner[f] < 1 || md->owner[f] > n || (f < F && (md->neighbour[f] < 1 || md->neighbour[f] > n))) {
      fcp_set_error("face %d: owner/neighbour out of range", f + 1);
      return FCP_EINVAL;
    }
  }

  // ---- periodic pairs (geometry.f90:251-257): face i of a periodic patch pairs with face i of its twin patch ----
  // pcell/pface/pord per boundary face; plist = the periodic-side boundary-face ordinals in patch order (the `l` of sparse_matrix.f90:262-293)
  std::vector<int32_t> pcell(std::max(B, 1), -1), pface(std::max(B, 1), -1), pord(std::max(B, 1), -1), plist;
  for (int32_t ib = 0; ib < c->nb; ++ib) {
    if (c->bctype[ib] != FCP_BC_PERIODIC) continue;
    if (c->nfaces[ib] == 0) continue;             // a partition that does not touch this periodic boundary
    const int32_t st = md->startFaceTwin ? md->startFaceTwin[ib] : -1;
    int32_t it = -1;                              // (empty patches share their startFace with the next one: match the size and the type too)
    for (int32_t jb = 0; jb < c->nb && it < 0; ++jb)
      if (jb != ib && c->startFace[jb] == st && c->nfaces[jb] == c->nfaces[ib] && c->bctype[jb] == FCP_BC_EMPTY) it = jb;
    if (it < 0) {
      fcp_set_error("periodic patch %d: startFaceTwin must name an 'empty' patch with the same number of faces", ib);
      return FCP_EINVAL;
    }
    for (int32_t i = 0; i < c->nfaces[ib]; ++i) {
      const int32_t fp = c->startFace[ib] + i, ft = st + i;
      const int32_t p = md->owner[fp] - 1, q = md->owner[ft] - 1;
      if (p == q) { fcp_set_error("periodic patch %d: face %d pairs a cell with itself", ib, i + 1); return FCP_EINVAL; }
      pcell[fp - F] = q; pface[fp - F] = ft; pord[fp - F] = i;
      pcell[ft - F] = p; pface[ft - F] = fp; pord[ft - F] = i;
      plist.push_back(fp - F);
    }
  }
  c->nper = (int32_t)plist.size();

  // ---- create_CSR_matrix (sparse_matrix.f90:110-260): rows ascending, columns ascending, diagonal embedded ----
  std::vector<int32_t> ia(n + 1, 0), diag(n);
  {
    std::vector<int32_t> cnt(n, 1);
    for (int32_t f = 0; f < F; ++f) { cnt[md->owner[f] - 1]++; cnt[md->neighbour[f] - 1]++; }
    for (int32_t b : plist) { cnt[md->owner[F + b] - 1]++; cnt[pcell[b]]++; }          // :141-171 twin entries
    ia[0] = 1;
    for (int32_t i = 0; i < n; ++i) ia[i + 1] = ia[i] + cnt[i];
  }
  const int64_t nnz = (int64_t)ia[n] - 1;
  std::vector<int32_t> ja(nnz);
  {
    std::vector<int32_t> pos(n);
    for (int32_t i = 0; i < n; ++i) { pos[i] = ia[i] - 1; ja[pos[i]++] = i + 1; }
    for (int32_t f = 0; f < F; ++f) {
      int32_t p = md->owner[f] - 1, q = md->neighbour[f] - 1;
      ja[pos[p]++] = q + 1;
      ja[pos[q]++] = p + 1;
    }
    for (int32_t b : plist) {
      int32_t p = md->owner[F + b] - 1, q = pcell[b];
      ja[pos[p]++] = q + 1;
      ja[pos[q]++] = p + 1;
    }
    parallel_for(n, [&](int64_t b, int64_t e) {
      for (int64_t i = b; i < e; ++i) {
        std::sort(ja.begin() + (ia[i] - 1), ja.begin() + (ia[i + 1] - 1));
        for (int32_t k = ia[i]; k < ia[i + 1]; ++k)
          if (ja[k - 1] == (int32_t)i + 1) { diag[i] = k; break; }
      }
    });
  }
  if (c->nper) {   // a periodic pair must not duplicate an existing entry (the reference's csr_to_k would alias the two coefficients)
    for (int32_t i = 0; i < n; ++i)
      for (int32_t k = ia[i]; k < ia[i + 1] - 1; ++k)
        if (ja[k - 1] == ja[k]) { fcp_set_error("periodic pair joins cells %d and %d, which already share a face", i + 1, ja[k]); return FCP_EINVAL; }
  }
  c->h_kPN.resize((size_t)F + c->nper);
  c->h_kNP.resize((size_t)F + c->nper);
  for (int32_t l = 0; l < c->nper; ++l) {   // :262-293, periodic faces in patch order after the inner faces
    const int32_t p = md->owner[F + plist[l]], q = pcell[plist[l]] + 1;
    c->h_kPN[F + l] = (int32_t)(std::lower_bound(ja.begin() + (ia[p - 1] - 1), ja.begin() + (ia[p] - 1), q) - ja.begin()) + 1;
    c->h_kNP[F + l] = (int32_t)(std::lower_bound(ja.begin() + (ia[q - 1] - 1), ja.begin() + (ia[q] - 1), p) - ja.begin()) + 1;
  }
  parallel_for(F, [&](int64_t b, int64_t e) {
    for (int64_t f = b; f < e; ++f) {   // csr_to_k, utils.f90:96-147 (first match in the row)
      int32_t p = md->owner[f], q = md->neighbour[f];
      c->h_kPN[f] = (int32_t)(std::lower_bound(ja.begin() + (ia[p - 1] - 1), ja.begin() + (ia[p] - 1), q) - ja.begin()) + 1;
      c->h_kNP[f] = (int32_t)(std::lower_bound(ja.begin() + (ia[q - 1] - 1), ja.begin() + (ia[q] - 1), p) - ja.begin()) + 1;
    }
  });

  // halo columns (src-par layout): one extra column per process face, the ghost slot of that face, appended to the
  // owner's row in boundary-face order (src-par/dpcg.f90:129-143 adds apr(ipro)*x(ghost) after the row sum)
  std::vector<std::vector<int32_t>> halo;
  std::vector<int32_t> halo_k(std::max(B, 1), -1);   // per boundary face: offset of its halo entry inside the owner's row
  if (c->npro) {
    halo.resize(n);
    for (int32_t i = 0; i < B; ++i)
      if (bft[i] == FCP_BC_PROCESS) {
        int32_t p = md->owner[F + i] - 1;
        halo_k[i] = (ia[p + 1] - ia[p]) + (int32_t)halo[p].size();
        halo[p].push_back(n + i);
      }
  }
  int rc = sell_from_csr(c->pat, n, c->npro ? c->nT : n, ia.data(), ja.data(), diag.data(), c->npro ? &halo : nullptr);
  if (rc != FCP_OK) { return rc; }
  std::vector<int64_t> slptr(c->pat.nslices + 1);
  FCP_CUDA(cudaMemcpy(slptr.data(), c->pat.slptr, slptr.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
  auto sellpos = [&](int32_t row0, int32_t off) -> int32_t { return (int32_t)(slptr[row0 >> 5] + (int64_t)off * 32 + (row0 & 31)); };
  if (c->npro) {
    std::vector<int32_t> aprpos;
    for (int32_t i = 0; i < B; ++i)
      if (bft[i] == FCP_BC_PROCESS) {
        aprpos.push_back(sellpos(md->owner[F + i] - 1, halo_k[i]));
        c->h_procface.push_back(F + i);
      }
    FCP_TRY(dev_upload(&c->d_aprpos, aprpos.data(), aprpos.size()));
    FCP_TRY(dev_upload(&c->d_procface, c->h_procface.data(), c->h_procface.size()));
  }

  // face -> SELL slot maps
  std::vector<int32_t> kPN(std::max(F, 1)), kNP(std::max(F, 1));
  parallel_for(F, [&](int64_t b, int64_t e) {
    for (int64_t f = b; f < e; ++f) {
      int32_t p = md->owner[f] - 1, q = md->neighbour[f] - 1;
      kPN[f] = sellpos(p, c->h_kPN[f] - ia[p]);
      kNP[f] = sellpos(q, c->h_kNP[f] - ia[q]);
    }
  });

  if (c->nper) {   // SELL positions of a(owner(face), cell across the pair) for both faces of every pair
    std::vector<int32_t> pslot(B, -1);
    for (int32_t l = 0; l < c->nper; ++l) {
      const int32_t b = plist[l], bt = pface[b] - F;
      const int32_t p = md->owner[F + b] - 1, q = pcell[b];
      pslot[b] = sellpos(p, c->h_kPN[F + l] - ia[p]);
      pslot[bt] = sellpos(q, c->h_kNP[F + l] - ia[q]);
    }
    FCP_TRY(dev_upload(&c->per_cell, pcell.data(), (size_t)B));
    FCP_TRY(dev_upload(&c->per_face, pface.data(), (size_t)B));
    FCP_TRY(dev_upload(&c->per_slot, pslot.data(), (size_t)B));
    std::vector<double> pdf(B, 0.0);              // quirk Q21: "Df(i)" = the Df of inner face number i, unless the host supplies the global mesh's value
    for (int32_t b = 0; b < B; ++b)
      if (pord[b] >= 0) {
        if (md->DfPeriodic) pdf[b] = md->DfPeriodic[b];
        else if (pord[b] < F) pdf[b] = md->Df[pord[b]];
        else { fcp_set_error("periodic patch with more faces (%d) than the mesh has inner faces: Df(i) of calcp_simple.f90:199 is out of range", pord[b] + 1); return FCP_EINVAL; }
      }
    FCP_TRY(dev_upload(&c->per_df, pdf.data(), (size_t)B));
  }

  // ---- cell -> face gather lists, faces in ascending face index -------------------------------------------------
  {
    std::vector<int32_t> cnt(n, 0), ptr(n + 1, 0);
    for (int32_t f = 0; f < F; ++f) { cnt[md->owner[f] - 1]++; cnt[md->neighbour[f] - 1]++; }
    for (int32_t f = F; f < c->nF; ++f) cnt[md->owner[f] - 1]++;
    const int32_t nsl = (n + 31) / 32;
    std::vector<int64_t> fsl(nsl + 1, 0);
    for (int32_t s = 0; s < nsl; ++s) {
      int32_t w = 0;
      for (int32_t r = s * 32; r < std::min(n, s * 32 + 32); ++r) w = std::max(w, cnt[r]);
      fsl[s + 1] = fsl[s] + (int64_t)w * 32;
    }
    const int64_t np = fsl[nsl];
    if (np >= (int64_t)2147483647) { fcp_set_error("face lists too large"); return FCP_EINVAL; }
    std::vector<int32_t> ent(np, 0), other(np, 0), slot(np, -1), fill(n, 0);
    // owner-ordered positions of the faces (fcp_ctx::og): SELL layout of the owner cells, faces of a cell in ascending index
    std::vector<int32_t> ocnt(n, 0), ofill(n, 0), gpos(std::max(c->nF, 1), 0), gent(np, 0);
    for (int32_t f = 0; f < c->nF; ++f) ocnt[md->owner[f] - 1]++;
    std::vector<int64_t> osl(nsl + 1, 0);
    for (int32_t s = 0; s < nsl; ++s) {
      int32_t w = 0;
      for (int32_t r = s * 32; r < std::min(n, s * 32 + 32); ++r) w = std::max(w, ocnt[r]);
      osl[s + 1] = osl[s] + (int64_t)w * 32;
    }
    if (osl[nsl] >= (int64_t)2147483647) { fcp_set_error("face lists too large"); return FCP_EINVAL; }
    c->og_n = std::max<int64_t>(osl[nsl], 1);
    for (int32_t f = 0; f < c->nF; ++f) {
      int32_t p = md->owner[f] - 1;
      int64_t pos = fsl[p >> 5] + (int64_t)fill[p]++ * 32 + (p & 31);
      const int32_t gp = (int32_t)(osl[p >> 5] + (int64_t)ofill[p]++ * 32 + (p & 31));
      gpos[f] = gp;
      if (f < F) {
        int32_t q = md->neighbour[f] - 1;
        ent[pos] = f + 1; other[pos] = q; slot[pos] = kPN[f]; gent[pos] = gp + 1;
        int64_t pos2 = fsl[q >> 5] + (int64_t)fill[q]++ * 32 + (q & 31);
        ent[pos2] = -(f + 1); other[pos2] = p; slot[pos2] = kNP[f]; gent[pos2] = -(gp + 1);
      } else {
        int32_t i = f - F;
        ent[pos] = f + 1; other[pos] = n + i; gent[pos] = gp + 1;
        slot[pos] = bft[i] == FCP_BC_PROCESS ? sellpos(p, halo_k[i]) : -1 - bft[i];
      }
    }
    c->fl.nnzp = np;
    for (int32_t i = 0; i < n; ++i) c->max_cell_faces = std::max(c->max_cell_faces, cnt[i]);
    FCP_TRY(dev_upload(&c->fl.slptr, fsl.data(), fsl.size()));
    FCP_TRY(dev_upload(&c->fl.len, cnt.data(), cnt.size()));
    FCP_TRY(dev_upload(&c->fl.ent, ent.data(), ent.size()));
    FCP_TRY(dev_upload(&c->fl.other, other.data(), other.size()));
    FCP_TRY(dev_upload(&c->fl.slot, slot.data(), slot.size()));
    FCP_TRY(dev_upload(&c->fl.gent, gent.data(), gent.size()));
    FCP_TRY(dev_upload(&c->d_gpos, gpos.data(), gpos.size()));
    std::vector<unsigned long long> kinds(n, 0ull);
    parallel_for(n, [&](int64_t b, int64_t e) {
      for (int64_t i = b; i < e; ++i) {
        unsigned long long w = 0;
        if (cnt[i] > 14) {
          w = 255ull << 56;
        } else {
          for (int32_t k = 0; k < cnt[i]; ++k) {
            const int32_t sl = slot[fsl[i >> 5] + (int64_t)k * 32 + (i & 31)];
            if (sl < 0) w |= (unsigned long long)(-sl & 15) << (4 * k);
          }
          w |= (unsigned long long)cnt[i] << 56;
        }
        kinds[i] = w;
      }
    });
    FCP_TRY(dev_upload(&c->fl.kinds, kinds.data(), kinds.size()));
  }

  // ---- mesh arrays --------------------------------------------------------------------------------------------
  {
    std::vector<int32_t> o0(c->nF), n0(std::max(F, 1));
    for (int32_t f = 0; f < c->nF; ++f) o0[f] = md->owner[f] - 1;
    for (int32_t f = 0; f < F; ++f) n0[f] = md->neighbour[f] - 1;
    FCP_TRY(dev_upload(&c->owner, o0.data(), o0.size()));
    FCP_TRY(dev_upload(&c->neigh, n0.data(), (size_t)F));
  }
  FCP_TRY(dev_upload(&c->arx, md->arx, (size_t)c->nF));
  FCP_TRY(dev_upload(&c->ary, md->ary, (size_t)c->nF));
  FCP_TRY(dev_upload(&c->arz, md->arz, (size_t)c->nF));
  FCP_TRY(dev_upload(&c->xf, md->xf, (size_t)c->nF));
  FCP_TRY(dev_upload(&c->yf, md->yf, (size_t)c->nF));
  FCP_TRY(dev_upload(&c->zf, md->zf, (size_t)c->nF));
  {
    // facint / Df are sized numFaces so that process faces can carry their own values (filled by fcp_comm_init)
    std::vector<double> tmp(c->nF, 0.5);
    std::copy(md->facint, md->facint + F, tmp.begin());
    FCP_TRY(dev_upload(&c->facint, tmp.data(), tmp.size()));
    std::fill(tmp.begin(), tmp.end(), 0.0);
    std::copy(md->Df, md->Df + F, tmp.begin());
    FCP_TRY(dev_upload(&c->Df, tmp.data(), tmp.size()));
  }
  {
    std::vector<double> tmp(c->nT, 0.0);
    const double *src[4] = {md->xc, md->yc, md->zc, md->vol};
    double **dst[4] = {&c->xc, &c->yc, &c->zc, &c->vol};
    for (int k = 0; k < 4; ++k) {
      std::copy(src[k], src[k] + n, tmp.begin());
      FCP_TRY(dev_upload(dst[k], tmp.data(), tmp.size()));
    }
  }
  FCP_TRY(dev_upload(&c->bftype, bft.data(), (size_t)std::max(B, 1)));
  FCP_TRY(dev_upload(&c->kPN, kPN.data(), (size_t)std::max(F, 1)));
  FCP_TRY(dev_upload(&c->kNP, kNP.data(), (size_t)std::max(F, 1)));
  return FCP_OK;
}
extern "C" int fcp_ctx_create(const fcp_mesh_desc *md, int device, fcp_ctx **out) {
  if (!md || !out) { fcp_set_error("fcp_ctx_create: null argument"); return FCP_EINVAL; }
  *out = nullptr;
  FCP_TRY(select_device(device));
  fcp_ctx *c = new fcp_ctx();
  const int rc = ctx_build(c, md, device);
  if (rc != FCP_OK) { fcp_ctx_destroy(c); return rc; }
  *out = c;
  return FCP_OK;
}

extern "C" int fcp_ctx_destroy(fcp_ctx *c) {
  if (!c) return FCP_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->comm) comm_free(c->comm);
  cudaFree(c->owner); cudaFree(c->neigh); cudaFree(c->arx); cudaFree(c->ary); cudaFree(c->arz);
  cudaFree(c->xf); cudaFree(c->yf); cudaFree(c->zf); cudaFree(c->facint); cudaFree(c->Df);
  cudaFree(c->xc); cudaFree(c->yc); cudaFree(c->zc); cudaFree(c->vol); cudaFree(c->bftype);
  cudaFree(c->kPN); cudaFree(c->kNP);
  cudaFree(c->fl.slptr); cudaFree(c->fl.len); cudaFree(c->fl.ent); cudaFree(c->fl.other); cudaFree(c->fl.slot); cudaFree(c->fl.kinds); cudaFree(c->fl.gent); cudaFree(c->d_gpos); cudaFree(c->og);
  for (int i = 0; i < FCP_F_COUNT; ++i) cudaFree(c->field[i]);
  for (int i = 0; i < 4; ++i) cudaFree(c->Dmat[i]);
  cudaFree(c->flushbuf); cudaFree(c->d_mmpart); cudaFree(c->d_sum);
  cudaFree(c->d_oface); cudaFree(c->d_flowo); cudaFree(c->d_csr_stage); cudaFree(c->d_aprpos); cudaFree(c->d_procface); cudaFree(c->d_proc_flip); cudaFree(c->d_ppref);
  cudaFree(c->per_cell); cudaFree(c->per_face); cudaFree(c->per_slot); cudaFree(c->per_df);
  sell_free(c->pat);
  krylov_ws_free(c->ws);
  if (c->t0) cudaEventDestroy(c->t0);
  if (c->t1) cudaEventDestroy(c->t1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return FCP_OK;
}

extern "C" int fcp_ctx_sizes(const fcp_ctx *c, int32_t *numCells, int32_t *numTotal, int32_t *numFaces, int32_t *nnz, int32_t *npro) {
  if (!c) return FCP_EINVAL;
  if (numCells) *numCells = c->n;
  if (numTotal) *numTotal = c->nT;
  if (numFaces) *numFaces = c->nF;
  if (nnz) *nnz = (int32_t)c->pat.nnz;
  if (npro) *npro = c->npro;
  return FCP_OK;
}

extern "C" int fcp_csr_pattern(const fcp_ctx *c, int32_t *ia, int32_t *ja, int32_t *diag, int32_t *icell_jcell, int32_t *jcell_icell) {
  if (!c) return FCP_EINVAL;
  if (ia) std::copy(c->pat.h_ia.begin(), c->pat.h_ia.end(), ia);
  if (ja) std::copy(c->pat.h_ja.begin(), c->pat.h_ja.end(), ja);
  if (diag) std::copy(c->pat.h_diag.begin(), c->pat.h_diag.end(), diag);
  if (icell_jcell) std::copy(c->h_kPN.begin(), c->h_kPN.end(), icell_jcell);
  if (jcell_icell) std::copy(c->h_kNP.begin(), c->h_kNP.end(), jcell_icell);
  return FCP_OK;
}

extern "C" int fcp_sync(fcp_ctx *c) {
  if (!c) return FCP_EINVAL;
  FCP_CUDA(cudaStreamSynchronize(c->stream));
  return FCP_OK;
}

// ---------------------------------------------------------------------------------------------
// fields
// ---------------------------------------------------------------------------------------------
// apr <-> halo entries of the SELL matrix
__global__ void k_apr_scatter(int32_t npro, const int32_t *__restrict__ pos, const double *__restrict__ apr, double *__restrict__ a) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npro) a[pos[i]] = apr[i];
}
__global__ void k_apr_gather(int32_t npro, const int32_t *__restrict__ pos, const double *__restrict__ a, double *__restrict__ apr) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npro) apr[i] = a[pos[i]];
}

extern "C" int fcp_field_upload(fcp_ctx *ctx, int field, const double *host, int64_t count) {
  if (!ctx || !host) return FCP_EINVAL;
  FCP_CUDA(cudaSetDevice(ctx->device));
  FIELD(d, field);
  int64_t cap = 0;
  FCP_TRY(field_count(ctx, field, &cap));
  if (field == FCP_F_A || field == FCP_F_H) {
    if (count != ctx->pat.nnz) { fcp_set_error("upload of a(nnz): count %lld != nnz %lld", (long long)count, (long long)ctx->pat.nnz); return FCP_EINVAL; }
    if (!ctx->d_csr_stage) FCP_TRY(dev_alloc(&ctx->d_csr_stage, (size_t)count));      // CSR-order staging of a(nnz), kept for the life of the context
    double *stage = ctx->d_csr_stage;
    FCP_CUDA(cudaMemcpyAsync(stage, host, sizeof(double) * (size_t)count, cudaMemcpyHostToDevice, ctx->stream));
    int rc = sell_values_from_csr(ctx->pat, stage, d, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    return rc;
  }
  if (count < 0 || count > cap) { fcp_set_error("field %d: count %lld exceeds extent %lld", field, (long long)count, (long long)cap); return FCP_EINVAL; }
  FCP_CUDA(cudaMemcpyAsync(d, host, sizeof(double) * (size_t)count, cudaMemcpyHostToDevice, ctx->stream));
  if (field == FCP_F_APR && ctx->npro) {
    FIELD(a, FCP_F_A);
    k_apr_scatter<<<(ctx->npro + 255) / 256, 256, 0, ctx->stream>>>(ctx->npro, ctx->d_aprpos, d, a);
    FCP_LAUNCHED();
  }
  FCP_CUDA(cudaStreamSynchronize(ctx->stream));
  return FCP_OK;
}

extern "C" int fcp_field_download(fcp_ctx *ctx, int field, double *host, int64_t count) {
  if (!ctx || !host) return FCP_EINVAL;
  FCP_CUDA(cudaSetDevice(ctx->device));
  FIELD(d, field);
  int64_t cap = 0;
  FCP_TRY(field_count(ctx, field, &cap));
  if (field == FCP_F_A || field == FCP_F_H) {
    if (count != ctx->pat.nnz) { fcp_set_error("download of a(nnz): count %lld != nnz %lld", (long long)count, (long long)ctx->pat.nnz); return FCP_EINVAL; }
    if (!ctx->d_csr_stage) FCP_TRY(dev_alloc(&ctx->d_csr_stage, (size_t)count));
    double *stage = ctx->d_csr_stage;
    int rc = sell_values_to_csr(ctx->pat, d, stage, ctx->stream);
    if (rc == FCP_OK) {
      cudaError_t e = cudaMemcpyAsync(host, stage, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream);
      if (e != cudaSuccess) rc = FCP_ECUDA;
    }
    cudaStreamSynchronize(ctx->stream);
    return rc;
  }
  if (count < 0 || count > cap) { fcp_set_error("field %d: count %lld exceeds extent %lld", field, (long long)count, (long long)cap); return FCP_EINVAL; }
  if (field == FCP_F_APR && ctx->npro) {
    FIELD(a, FCP_F_A);
    k_apr_gather<<<(ctx->npro + 255) / 256, 256, 0, ctx->stream>>>(ctx->npro, ctx->d_aprpos, a, d);
    FCP_LAUNCHED();
  }
  FCP_CUDA(cudaMemcpyAsync(host, d, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
  FCP_CUDA(cudaStreamSynchronize(ctx->stream));
  return FCP_OK;
}

__global__ void k_fill(int64_t n, double *x, double v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = v;
}
extern "C" int fcp_field_fill(fcp_ctx *ctx, int field, double value) {
  if (!ctx) return FCP_EINVAL;
  FIELD(d, field);
  int64_t cnt = 0;
  FCP_TRY(field_count(ctx, field, &cnt));
  if (cnt == 0) return FCP_OK;
  if (value == 0.0) {
    FCP_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * (size_t)cnt, ctx->stream));
  } else {
    k_fill<<<(int)std::min<int64_t>((cnt + 255) / 256, 148 * 16), 256, 0, ctx->stream>>>(cnt, d, value);
    FCP_LAUNCHED();
    FCP_CHECK_LAUNCH();
  }
  return FCP_OK;
}
extern "C" int fcp_field_copy(fcp_ctx *ctx, int dst, int src) {
  if (!ctx) return FCP_EINVAL;
  FIELD(d, dst);
  FIELD(s, src);
  int64_t cd = 0, cs = 0;
  FCP_TRY(field_count(ctx, dst, &cd));
  FCP_TRY(field_count(ctx, src, &cs));
  if (cd != cs) { fcp_set_error("fcp_field_copy: extents differ"); return FCP_EINVAL; }
  FCP_CUDA(cudaMemcpyAsync(d, s, sizeof(double) * (size_t)cd, cudaMemcpyDeviceToDevice, ctx->stream));
  return FCP_OK;
}
// dst = alpha*x + beta*y, elementwise over the common extent: the streaming part of the fvEquation operators
// (fvImplicit/fvEquation.f90:158-404: operator(+), operator(-), operator(==) on coef(nnz) and source(numCells))
__global__ void __launch_bounds__(256) k_axpby(int64_t n, double alpha, const double *x, double beta, const double *y, double *dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double xv = x[i], yv = y[i];
    dst[i] = (alpha == 1.0 ? xv : alpha * xv) + (beta == 1.0 ? yv : beta == -1.0 ? -yv : beta * yv);
  }
}
extern "C" int fcp_field_axpby(fcp_ctx *ctx, int dst_field, double alpha, int x_field, double beta, int y_field) {
  if (!ctx) return FCP_EINVAL;
  FIELD(d, dst_field);
  FIELD(x, x_field);
  FIELD(y, y_field);
  int64_t cd = 0, cx = 0, cy = 0;
  FCP_TRY(field_count(ctx, dst_field, &cd));
  FCP_TRY(field_count(ctx, x_field, &cx));
  FCP_TRY(field_count(ctx, y_field, &cy));
  if (cd != cx || cd != cy) { fcp_set_error("fcp_field_axpby: extents differ"); return FCP_EINVAL; }
  if (cd == 0) return FCP_OK;
  k_axpby<<<(int)std::min<int64_t>((cd + 255) / 256, 148 * 16), 256, 0, ctx->stream>>>(cd, alpha, x, beta, y, d);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
extern "C" int fcp_field_devptr(fcp_ctx *ctx, int field, void **devptr, int64_t *count) {
  if (!ctx || !devptr) return FCP_EINVAL;
  FIELD(d, field);
  *devptr = d;
  if (count) FCP_TRY(field_count(ctx, field, count));
  return FCP_OK;
}

// ---------------------------------------------------------------------------------------------
// operators
// ---------------------------------------------------------------------------------------------
extern "C" int fcp_spmv(fcp_ctx *ctx, int x_field, int y_field) {
  if (!ctx) return FCP_EINVAL;
  FIELD(x, x_field);
  FIELD(y, y_field);
  FIELD(a, FCP_F_A);
  if (ctx->comm) FCP_TRY(comm_exchange(ctx, x, 1));
  int rc;
  FCP_PROF(&ctx->prof, FCP_K_SPMV, ctx->stream, rc = sell_spmv(ctx->pat, a, x, y, ctx->stream));
  return rc;
}

extern "C" int fcp_csrsolve(fcp_ctx *ctx, int solver, int fi_field, int rhs_field, int32_t itr_max, double tol_abs, double tol_rel,
                            fcp_report *rep) {
  if (!ctx) return FCP_EINVAL;
  FCP_CUDA(cudaSetDevice(ctx->device));
  FIELD(fi, fi_field);
  FIELD(rhs, rhs_field);
  FIELD(a, FCP_F_A);
  return krylov_solve(solver, ctx->pat, a, fi, rhs, ctx->ws, itr_max, tol_abs, tol_rel, rep, ctx->stream, ctx->comm, ctx);
}

extern "C" int fcp_report_line(const fcp_report *rep, const char *chvar, char *buf, int buflen) {
  if (!rep || !buf) return FCP_EINVAL;
  // linear_solvers.f90:354-355 (dpcg), :540-541 (iccg), :781-782 (bicgstab); early return :267-268
  const char *name = rep->solver == FCP_SOLVER_DPCG ? "PCG(Jacobi)" : rep->solver == FCP_SOLVER_ICCG ? "PCG(IC0)" :
                     rep->solver == FCP_SOLVER_GAUSS_SEIDEL ? "Gauss-Seidel" : "BiCGStab(ILU(0))";
  auto e103 = [](double v, char *out) {   // Fortran 1PE10.3
    char t[64];
    snprintf(t, sizeof(t), "%10.3E", v);
    strcpy(out, t);
  };
  char r0[64], r1[64];
  if (rep->solver == FCP_SOLVER_GAUSS_SEIDEL && rep->iters == 1 && rep->factor == 0.0) {   // :151-157: the early return comes after the first sweep
    e103(rep->res0, r0);
    snprintf(buf, buflen, "  %s:  Solving for %s, Initial residual = %s, Final residual = %s, No Iterations 1", name, chvar ? chvar : "", r0, r0);
  } else if (rep->iters == 0 && rep->factor == 0.0) {
    e103(rep->res0, r0);
    snprintf(buf, buflen, "  %s:  Solving for %s, Initial residual = %s, Final residual = %s, No Iterations 0", name, chvar ? chvar : "", r0, r0);
  } else {
    e103(rep->resor, r0);
    e103(rep->resl / rep->factor, r1);
    snprintf(buf, buflen, "  %s:  Solving for %s, Initial residual = %s, Final residual = %s, No Iterations %d", name, chvar ? chvar : "", r0, r1,
             rep->iters);
  }
  return FCP_OK;
}

extern "C" int fcp_create_lsq_grad_matrix(fcp_ctx *ctx, int method) {
  if (!ctx) return FCP_EINVAL;
  if (method != FCP_GRAD_LSQ && method != FCP_GRAD_LSQ_DM && method != FCP_GRAD_LSQ_QR) {
    fcp_set_error("create_lsq_grad_matrix: method %d is not a least-squares method", method);
    return FCP_EINVAL;
  }
  if (method == FCP_GRAD_LSQ_QR) {
    if (ctx->max_cell_faces > 6) {   // gradients.f90:924  m=6: D(3,6,numCells)
      fcp_set_error("create_matrix_lsq_qr holds at most 6 faces per cell (gradients.f90:924); this mesh has a cell with %d", ctx->max_cell_faces);
      return FCP_EINVAL;
    }
    if (ctx->comm) { fcp_set_error("the QR gradient has no src-par twin; single-GPU only"); return FCP_ESTATE; }
    if (!ctx->Dmat[method]) FCP_TRY(dev_alloc(&ctx->Dmat[method], (size_t)18 * ctx->n));
    return fvm_lsq_qr_matrix(ctx, ctx->Dmat[method]);
  }
  if (!ctx->Dmat[method]) FCP_TRY(dev_alloc(&ctx->Dmat[method], (size_t)9 * ctx->n));
  return fvm_lsq_matrix(ctx, method == FCP_GRAD_LSQ_DM, ctx->Dmat[method]);
}

extern "C" int fcp_grad(fcp_ctx *ctx, int method, int phi_field, int grad_field, int lsq_row2_reference) {
  if (!ctx) return FCP_EINVAL;
  if (!fcp_is_gradient_field(grad_field)) { fcp_set_error("fcp_grad: field %d is not a gradient field", grad_field); return FCP_EINVAL; }
  FIELD(phi, phi_field);
  FIELD(g, grad_field);
  if (ctx->comm) FCP_TRY(comm_exchange(ctx, phi, 1));              // src-par/gradients.f90:123
  int rc;
  if (method == FCP_GRAD_GAUSS) rc = fvm_grad_gauss(ctx, phi, g);
  else if (method == FCP_GRAD_LSQ || method == FCP_GRAD_LSQ_DM) {
    if (!ctx->Dmat[method]) { fcp_set_error("fcp_grad: call fcp_create_lsq_grad_matrix(method=%d) first", method); return FCP_ESTATE; }
    rc = fvm_grad_lsq(ctx, method == FCP_GRAD_LSQ_DM, ctx->Dmat[method], phi, g, lsq_row2_reference);
  } else if (method == FCP_GRAD_LSQ_QR) {
    if (!ctx->Dmat[method]) { fcp_set_error("fcp_grad: call fcp_create_lsq_grad_matrix(method=%d) first", method); return FCP_ESTATE; }
    rc = fvm_grad_lsq_qr(ctx, ctx->Dmat[method], phi, g);
  } else { fcp_set_error("fcp_grad: unknown method %d", method); return FCP_EINVAL; }
  FCP_TRY(rc);
  if (ctx->comm) FCP_TRY(comm_exchange(ctx, g, 3));                // src-par/gradients.f90:168-170
  return FCP_OK;
}

extern "C" int fcp_slope_limiter(fcp_ctx *ctx, int limiter, int phi_field, int grad_field) {
  if (!ctx) return FCP_EINVAL;
  if (limiter < FCP_LIMITER_NONE || limiter > FCP_LIMITER_MULTIDIMENSIONAL) { fcp_set_error("unknown limiter %d", limiter); return FCP_EINVAL; }
  if (!fcp_is_gradient_field(grad_field)) { fcp_set_error("fcp_slope_limiter: field %d is not a gradient field", grad_field); return FCP_EINVAL; }
  FIELD(phi, phi_field);
  FIELD(g, grad_field);
  FCP_TRY(fvm_slope_limiter(ctx, limiter, phi, g));
  if (ctx->comm && limiter != FCP_LIMITER_NONE) FCP_TRY(comm_exchange(ctx, g, 3));
  return FCP_OK;
}

extern "C" int fcp_grad_opt(fcp_ctx *ctx, int method, int limiter, int phi_field, int grad_field) {
  // gradients.f90:217-278; dPhidxi = 0 first (:238) -- every method below overwrites all cell entries and zeroes the
  // boundary slots, which is the same state
  FCP_TRY(fcp_grad(ctx, method, phi_field, grad_field, 1));
  return fcp_slope_limiter(ctx, limiter, phi_field, grad_field);
}

extern "C" int fcp_laplacian(fcp_ctx *ctx, int mu_field, int phi_field) {
  if (!ctx) return FCP_EINVAL;
  FIELD(mu, mu_field);
  FIELD(phi, phi_field);
  FIELD(a, FCP_F_A);
  FIELD(su, FCP_F_SU);
  if (ctx->comm) { FCP_TRY(comm_exchange(ctx, mu, 1)); }
  return fvm_laplacian(ctx, mu, phi, a, su);
}

static int gradp_impl(fcp_ctx *ctx, int pscheme, double *p, const CorrectArgs *ca) {
  FIELD(apu, FCP_F_APU);
  FIELD(su, FCP_F_SU);
  FIELD(sv, FCP_F_SV);
  FIELD(sw, FCP_F_SW);
  FIELD(g, FCP_F_DPDXI);
  double *gtmp = nullptr;
  if (pscheme == FCP_PSCHEME_CENTRAL) { FIELD(t, FCP_F_G1); gtmp = t; }
  if (ctx->comm) FCP_TRY(comm_exchange(ctx, p, 1));
  return fvm_gradp(ctx, pscheme, p, apu, su, sv, sw, g, gtmp, ca);
}
extern "C" int fcp_gradp_and_sources(fcp_ctx *ctx, int pscheme, int p_field) {
  if (!ctx) return FCP_EINVAL;
  if (pscheme < 0 || pscheme > 2) { fcp_set_error("unknown pscheme %d", pscheme); return FCP_EINVAL; }   // nablap.f90:106-110 stops
  FIELD(p, p_field);
  return gradp_impl(ctx, pscheme, p, nullptr);
}

static int ensure_outlet_list(fcp_ctx *c) {
  if (c->d_oface) return FCP_OK;        // (a rank without outlet faces keeps an empty list and still joins the global sum)
  std::vector<int32_t> of;
  for (int32_t ib = 0; ib < c->nb; ++ib)
    if (c->bctype[ib] == FCP_BC_OUTLET)
      for (int32_t i = 0; i < c->nfaces[ib]; ++i) of.push_back(c->startFace[ib] + i);
  c->nout = (int32_t)of.size();
  if (of.empty()) of.push_back(0);
  return dev_upload(&c->d_oface, of.data(), of.size());
}

extern "C" int fcp_assemble_pcorr_simple(fcp_ctx *ctx, int const_mflux, double flomas) {
  if (!ctx) return FCP_EINVAL;
  FIELD(den, FCP_F_DEN); FIELD(u, FCP_F_U); FIELD(v, FCP_F_V); FIELD(w, FCP_F_W); FIELD(p, FCP_F_P); FIELD(pp, FCP_F_PP);
  FIELD(g, FCP_F_DPDXI); FIELD(apu, FCP_F_APU); FIELD(a, FCP_F_A); FIELD(su, FCP_F_SU); FIELD(fl, FCP_F_FLMASS);
  const double *apv = nullptr, *apw = nullptr;
  if (ctx->nper) { FIELD(x1, FCP_F_APV); FIELD(x2, FCP_F_APW); apv = x1; apw = x2; }   // facefluxmass2_periodic weights
  if (ctx->flux_variant == 1) {
    // quirk Q10: the MPI tree's facefluxmass on inner faces.  src-par/calcp_simple.f90:40-42: "tentative velocity gradients used for velocity
    // interpolation: call grad(U,dUdxi) ..." -- with the gradient method the host selected, and BEFORE adjustMassFlow rescales the outlet values (:113)
    FCP_TRY(fcp_grad(ctx, ctx->flux_grad_method, FCP_F_U, FCP_F_DUDXI, 1));
    FCP_TRY(fcp_grad(ctx, ctx->flux_grad_method, FCP_F_V, FCP_F_DVDXI, 1));
    FCP_TRY(fcp_grad(ctx, ctx->flux_grad_method, FCP_F_W, FCP_F_DWDXI, 1));
  }
  if (!const_mflux && ctx->g_outlet) {                             // adjustMassFlow, calcp_simple.f90:125
    FCP_TRY(ensure_outlet_list(ctx));
    FCP_TRY(fvm_adjust_mass_flow(ctx, ctx->nout, ctx->d_oface, den, u, v, w, fl, flomas));
  }
  if (ctx->comm) {   // ghost values of everything facefluxmass reads on the far side of a process face
    double *sc[] = {den, u, v, w, p, apu};
    for (double *x : sc) FCP_TRY(comm_exchange(ctx, x, 1));
    FCP_TRY(comm_exchange(ctx, g, 3));
  }
  AsmArgs args{den, u, v, w, p, g, apu, apv, apw, pp, u, v, w, a, su, fl};
  if (ctx->flux_variant == 1) {
    FIELD(gu, FCP_F_DUDXI); FIELD(gv, FCP_F_DVDXI); FIELD(gw, FCP_F_DWDXI); FIELD(x1, FCP_F_APV); FIELD(x2, FCP_F_APW);
    args.gU = gu; args.gV = gv; args.gW = gw; args.apv = x1; args.apw = x2;
  }
  return fvm_assemble_pcorr(ctx, args);
}

// which mass-flux routine the inner faces of the SIMPLE p' assembly use.  variant 0 (default): facefluxmass2 of the serial
// tree (calcp_simple.f90:90); variant 1: facefluxmass of the MPI tree (src-par/calcp_simple.f90:52), with the velocity gradients computed by
// `grad_method` (FCP_GRAD_*) as that routine's `call grad(U,dUdxi)` does.  calcp_piso is not affected (the MPI tree has no PISO).
extern "C" int fcp_set_flux_variant(fcp_ctx *ctx, int variant, int grad_method) {
  if (!ctx) return FCP_EINVAL;
  if (variant != 0 && variant != 1) { fcp_set_error("fcp_set_flux_variant: unknown variant %d", variant); return FCP_EINVAL; }
  if (variant == 1 && (grad_method < FCP_GRAD_GAUSS || grad_method > FCP_GRAD_LSQ_QR)) { fcp_set_error("fcp_set_flux_variant: unknown gradient method %d", grad_method); return FCP_EINVAL; }
  ctx->flux_variant = variant;
  ctx->flux_grad_method = grad_method;
  return FCP_OK;
}

// pp(pRefCell) on the rank that owns the reference cell, 0 elsewhere (pRefCell <= 0 = "not on this rank")
__global__ void k_pick_ref(const double *pp, int32_t pRefCell, double *out) { out[0] = pRefCell > 0 ? pp[pRefCell - 1] : 0.0; }
// process faces: flmass += apr*(pp(ghost) - pp(owner))       src-par/calcp_simple.f90:222-244
__global__ void k_correct_flux_proc(int32_t npro, const int32_t *__restrict__ pface, const int32_t *__restrict__ aprpos, const int32_t *__restrict__ owner,
                                    int32_t n, int32_t F, const double *__restrict__ a, const double *__restrict__ pp, double *__restrict__ flmass) {
  int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npro) return;
  const int32_t f = pface[i];
  flmass[f] = flmass[f] + a[aprpos[i]] * (pp[n + (f - F)] - pp[owner[f]]);
}

static int fvm_correct_flux_proc(fcp_ctx *ctx, const double *a, const double *pp, double *fl) {
  k_correct_flux_proc<<<(ctx->npro + 255) / 256, 256, 0, ctx->stream>>>(ctx->npro, ctx->d_procface, ctx->d_aprpos, ctx->owner, ctx->n, ctx->F, a, pp, fl);
  FCP_LAUNCHED();
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}

extern "C" int fcp_correct_simple(fcp_ctx *ctx, int pscheme, double urfp, int32_t pRefCell) {
  if (!ctx) return FCP_EINVAL;
  // multi-GPU: pRefCell is the LOCAL index on the rank that owns the reference cell and <= 0 on every other rank
  if (pRefCell > ctx->n || (pRefCell < 1 && !ctx->comm)) { fcp_set_error("pRefCell %d out of range", pRefCell); return FCP_EINVAL; }
  FIELD(den, FCP_F_DEN); FIELD(u, FCP_F_U); FIELD(v, FCP_F_V); FIELD(w, FCP_F_W); FIELD(p, FCP_F_P); FIELD(pp, FCP_F_PP);
  FIELD(apu, FCP_F_APU); FIELD(apv, FCP_F_APV); FIELD(apw, FCP_F_APW); FIELD(a, FCP_F_A); FIELD(fl, FCP_F_FLMASS);
  if (ctx->comm) FCP_TRY(comm_exchange(ctx, pp, 1));
  FCP_TRY(fvm_correct_flux(ctx, a, pp, fl));                                   // calcp_simple.f90:331-341
  if (ctx->npro) FCP_TRY(fvm_correct_flux_proc(ctx, a, pp, fl));
  if (ctx->nper) FCP_TRY(fvm_correct_flux_periodic(ctx, a, pp, fl));                                // :350-373
  if (ctx->has_pressure_patch) FCP_TRY(fvm_correct_pressure_bnd(ctx, den, apu, pp, u, v, w, fl));   // :345-391
  const double *ppref = ctx->g_pressure_patch ? nullptr : pp + (pRefCell - 1);                      // :399-407
  if (ctx->comm && !ctx->g_pressure_patch) {
    // the broadcast of ppref (src-par/calcp_simple.f90:195-198, quirk Q12) as a rank-ordered sum of {pp(ref), 0, 0, ...}
    if (!ctx->d_ppref) FCP_TRY(dev_alloc(&ctx->d_ppref, 4));
    k_pick_ref<<<1, 1, 0, ctx->stream>>>(pp, pRefCell, ctx->d_ppref);
    FCP_LAUNCHED();
    FCP_TRY(comm_allgather_sum(ctx->comm, ctx->d_ppref, 1, ctx->stream));
    ppref = ctx->d_ppref;
  }
  CorrectArgs ca{u, v, w, p, apu, apv, apw, urfp, ppref};
  return gradp_impl(ctx, pscheme, pp, &ca);                                    // :412-429
}

extern "C" int fcp_nonorth_corrector(fcp_ctx *ctx) {
  if (!ctx) return FCP_EINVAL;
  FIELD(den, FCP_F_DEN); FIELD(apu, FCP_F_APU); FIELD(g, FCP_F_DPDXI); FIELD(su, FCP_F_SU); FIELD(fl, FCP_F_FLMASS);
  if (ctx->comm) { FCP_TRY(comm_exchange(ctx, g, 3)); FCP_TRY(comm_exchange(ctx, den, 1)); FCP_TRY(comm_exchange(ctx, apu, 1)); }
  return fvm_nonorth(ctx, den, apu, g, su, fl);
}

extern "C" int fcp_calcp_simple(fcp_ctx *ctx, const fcp_simple_params *prm, fcp_report *rep) {
  if (!ctx || !prm) return FCP_EINVAL;
  FCP_TRY(fcp_assemble_pcorr_simple(ctx, prm->const_mflux, prm->flomas));
  for (int ipcorr = 1; ipcorr <= prm->npcor; ++ipcorr) {                       // calcp_simple.f90:318
    if (prm->zero_pp) FCP_TRY(fcp_field_fill(ctx, FCP_F_PP, 0.0));
    FCP_TRY(fcp_csrsolve(ctx, prm->solver, FCP_F_PP, FCP_F_SU, prm->maxiter, prm->tol_abs, prm->tol_rel, rep ? &rep[ipcorr - 1] : nullptr));
    FCP_TRY(fcp_correct_simple(ctx, prm->pscheme, prm->urfp, prm->pRefCell));
    if (ipcorr != prm->npcor) FCP_TRY(fcp_nonorth_corrector(ctx));            // :433-455
  }
  return FCP_OK;
}

// calcuvw   Velocity/velocity.f90:50-750
extern "C" int fcp_calcuvw(fcp_ctx *ctx, const fcp_uvw_params *prm, fcp_report *rep) {
  if (!ctx || !prm) return FCP_EINVAL;
  if (prm->cscheme < 0 || prm->cscheme >= FCP_CS_COUNT) { fcp_set_error("calcuvw: non-existing interpolation scheme %d", prm->cscheme); return FCP_EINVAL; }   // interpolation.f90:643-646 stops
  if (prm->tscheme < 0 || prm->tscheme > 3) { fcp_set_error("calcuvw: unknown time scheme %d (Crank-Nicolson is not built)", prm->tscheme); return FCP_EINVAL; }
  if (prm->tscheme && !(prm->timestep > 0.0)) { fcp_set_error("calcuvw: timestep must be positive"); return FCP_EINVAL; }
  for (int q = 0; q < 3; ++q) if (!(prm->urf[q] > 0.0)) { fcp_set_error("calcuvw: urfU(%d) must be positive", q + 1); return FCP_EINVAL; }
  FCP_CUDA(cudaSetDevice(ctx->device));
  FIELD(u, FCP_F_U); FIELD(v, FCP_F_V); FIELD(w, FCP_F_W); FIELD(p, FCP_F_P); FIELD(den, FCP_F_DEN); FIELD(vis, FCP_F_VIS); FIELD(visw, FCP_F_VISW);
  FIELD(fl, FCP_F_FLMASS); FIELD(a, FCP_F_A); FIELD(su, FCP_F_SU); FIELD(sv, FCP_F_SV); FIELD(sw, FCP_F_SW);
  FIELD(spu, FCP_F_SPU); FIELD(spv, FCP_F_SPV); FIELD(sp, FCP_F_SP); FIELD(apu, FCP_F_APU); FIELD(apv, FCP_F_APV); FIELD(apw, FCP_F_APW);
  FIELD(gu, FCP_F_DUDXI); FIELD(gv, FCP_F_DVDXI); FIELD(gw, FCP_F_DWDXI);
  UvwArgs g{};
  g.u = u; g.v = v; g.w = w; g.vis = vis; g.visw = visw; g.den = den; g.flmass = fl; g.dUdxi = gu; g.dVdxi = gv; g.dWdxi = gw;
  g.a = a; g.su = su; g.sv = sv; g.sw = sw; g.spu = spu; g.spv = spv; g.sp = sp;
  g.gds = prm->gds; g.timestep = prm->timestep; g.gradPcmf = prm->gradPcmf; g.viscos = prm->viscos;
  g.cscheme = prm->cscheme; g.tscheme = prm->tscheme; g.const_mflux = prm->const_mflux;
  if (prm->tscheme >= 1) { FIELD(uo, FCP_F_UO); FIELD(vo, FCP_F_VO); FIELD(wo, FCP_F_WO); g.uo = uo; g.vo = vo; g.wo = wo; }
  if (prm->tscheme >= 2) { FIELD(uoo, FCP_F_UOO); FIELD(voo, FCP_F_VOO); FIELD(woo, FCP_F_WOO); g.uoo = uoo; g.voo = voo; g.woo = woo; }
  if (prm->tscheme >= 3) { FIELD(x1, FCP_F_UOOO); FIELD(x2, FCP_F_VOOO); FIELD(x3, FCP_F_WOOO); g.uooo = x1; g.vooo = x2; g.wooo = x3; }
  if (prm->piso) { FIELD(rU, FCP_F_RU); FIELD(rV, FCP_F_RV); FIELD(rW, FCP_F_RW); g.rU = rU; g.rV = rV; g.rW = rW; }
  FCP_TRY(fvm_update_vel_bnd(ctx, u, v, w));                                                   // :170
  if (prm->grad_method != FCP_GRAD_GAUSS && !ctx->Dmat[prm->grad_method]) FCP_TRY(fcp_create_lsq_grad_matrix(ctx, prm->grad_method));
  FCP_TRY(fcp_grad_opt(ctx, prm->grad_method, prm->limiter, FCP_F_U, FCP_F_DUDXI));            // :171-173
  FCP_TRY(fcp_grad_opt(ctx, prm->grad_method, prm->limiter, FCP_F_V, FCP_F_DVDXI));
  FCP_TRY(fcp_grad_opt(ctx, prm->grad_method, prm->limiter, FCP_F_W, FCP_F_DWDXI));
  FCP_TRY(fcp_gradp_and_sources(ctx, prm->pscheme, FCP_F_P));                                  // :177
  if (ctx->comm) {
    double *sc[] = {vis, den};
    for (double *x : sc) FCP_TRY(comm_exchange(ctx, x, 1));
  }
  FCP_TRY(fvm_uvw_assemble(ctx, g));                                                           // :187-568
  const int fields[3] = {FCP_F_U, FCP_F_V, FCP_F_W};
  double *phi[3] = {u, v, w}, *spq[3] = {spu, spv, sp}, *srcq[3] = {su, sv, sw}, *apq[3] = {apu, apv, apw};
  for (int q = 0; q < 3; ++q) {                                                                // :602-750
    FCP_TRY(fvm_uvw_diag(ctx, a, spq[q], srcq[q], phi[q], apq[q], su, prm->urf[q], q > 0));
    FCP_TRY(fcp_csrsolve(ctx, prm->solver, fields[q], FCP_F_SU, prm->maxiter, prm->tol_abs, prm->tol_rel, rep ? &rep[q] : nullptr));
  }
  return FCP_OK;
}

// constant_mass_flow_forcing   src/cappuccino/constant_mass_flow_forcing.f90 (called at calcp_simple.f90:468 / calcp_piso.f90:492)
extern "C" int fcp_constant_mass_flow_forcing(fcp_ctx *ctx, double magUbar, double *gradPcmf, double *magUbarStar) {
  if (!ctx || !gradPcmf) return FCP_EINVAL;
  FCP_CUDA(cudaSetDevice(ctx->device));
  FIELD(u, FCP_F_U); FIELD(apu, FCP_F_APU);
  if (!ctx->d_sum) FCP_TRY(dev_alloc(&ctx->d_sum, 8));
  FCP_TRY(fvm_cmf_forcing(ctx, magUbar, apu, u, ctx->d_sum));
  double h[5];
  FCP_CUDA(cudaMemcpyAsync(h, ctx->d_sum, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  FCP_CUDA(cudaStreamSynchronize(ctx->stream));
  *gradPcmf = *gradPcmf + h[4];                      // :33  gradPcmf = gradPcmf + gragPplus
  if (magUbarStar) *magUbarStar = h[3];
  return FCP_OK;
}
// updateBoundary(phi)   src/finiteVolume/boundary/updateBoundary.f90
extern "C" int fcp_update_boundary(fcp_ctx *ctx, int field) {
  if (!ctx) return FCP_EINVAL;
  if (field < 0 || field >= FCP_F_COUNT || (field >= FCP_F_DUDXI && field <= FCP_F_H) || fcp_is_gradient_field(field)) { fcp_set_error("fcp_update_boundary: field %d is not a scalar cell field", field); return FCP_EINVAL; }
  FCP_CUDA(cudaSetDevice(ctx->device));
  FIELD(phi, field);
  return fvm_update_boundary(ctx, phi);
}

// calcsc: the scalar transport template   TurbulenceModels/k_epsilon_rlzb.f90:52-790 + fluxes/scalar_fluxes.f90
extern "C" int fcp_calcsc(fcp_ctx *ctx, const fcp_scalar_params *prm, int phi_field, fcp_report *rep, double *fimin, double *fimax) {
  if (!ctx || !prm) return FCP_EINVAL;
  if (prm->kind < FCP_SC_GENERIC || prm->kind > FCP_SC_OMEGA_SST) { fcp_set_error("calcsc: unknown kind %d", prm->kind); return FCP_EINVAL; }
  if (prm->cscheme < 0 || prm->cscheme >= FCP_CS_COUNT) { fcp_set_error("calcsc: non-existing interpolation scheme %d", prm->cscheme); return FCP_EINVAL; }
  if (prm->tscheme < 0 || prm->tscheme > 2) { fcp_set_error("calcsc: unknown time scheme %d (Crank-Nicolson is not built)", prm->tscheme); return FCP_EINVAL; }
  if (prm->tscheme && !(prm->timestep > 0.0)) { fcp_set_error("calcsc: timestep must be positive"); return FCP_EINVAL; }
  if (!(prm->urf > 0.0)) { fcp_set_error("calcsc: urf must be positive"); return FCP_EINVAL; }
  const bool is_k = prm->kind == FCP_SC_TKE_RLZB || prm->kind == FCP_SC_TKE_SST, is_sst = prm->kind == FCP_SC_TKE_SST || prm->kind == FCP_SC_OMEGA_SST;
  if (prm->kind != FCP_SC_GENERIC && phi_field != (is_k ? FCP_F_TE : FCP_F_ED)) {
    fcp_set_error("calcsc: kind %d solves for field %d", prm->kind, is_k ? FCP_F_TE : FCP_F_ED);
    return FCP_EINVAL;
  }
  if (phi_field < 0 || phi_field >= FCP_F_COUNT || (phi_field >= FCP_F_DUDXI && phi_field <= FCP_F_H) || fcp_is_gradient_field(phi_field) || phi_field == FCP_F_SCTMP) {
    fcp_set_error("calcsc: field %d is not a scalar cell field", phi_field);
    return FCP_EINVAL;
  }
  if (ctx->comm && is_sst && ctx->npro && !ctx->d_proc_flip) {   // sigma is taken from the OWNER of a face; a process face sees its local cell as owner on both ranks
    fcp_set_error("calcsc: the SST pair on a partitioned mesh needs the orientation of the process faces: call fcp_set_process_orientation after fcp_comm_init");
    return FCP_ESTATE;
  }
  FCP_CUDA(cudaSetDevice(ctx->device));
  FIELD(phi, phi_field); FIELD(den, FCP_F_DEN); FIELD(vis, FCP_F_VIS); FIELD(fl, FCP_F_FLMASS);
  // the gradient of the scalar: G0, except for the SST pair, whose omega equation needs both grad k and grad omega (cross diffusion)
  const int gfield = prm->kind == FCP_SC_TKE_SST ? FCP_F_DTEDXI : prm->kind == FCP_SC_OMEGA_SST ? FCP_F_DEDDXI : FCP_F_G0;
  FIELD(a, FCP_F_A); FIELD(su, FCP_F_SU); FIELD(sp, FCP_F_SP); FIELD(g, gfield); FIELD(tmp, FCP_F_SCTMP);
  ScParams q{};
  q.lowre = prm->lowre;
  q.kind = prm->kind; q.cscheme = prm->cscheme; q.tscheme = prm->tscheme;
  q.gds = prm->gds; q.prtr = prm->prtr; q.viscos = prm->viscos; q.densit = prm->densit; q.timestep = prm->timestep; q.urf = prm->urf;
  q.phi = phi; q.den = den; q.vis = vis; q.flmass = fl; q.grad = g; q.a = a; q.su = su; q.sp = sp; q.phi_new = tmp; q.phi_out = phi;
  if (prm->tscheme >= 1) { FIELD(x, FCP_F_PHIO); q.phio = x; }
  if (prm->tscheme >= 2) { FIELD(x, FCP_F_PHIOO); q.phioo = x; }
  if (prm->kind == FCP_SC_GENERIC) { FIELD(x1, FCP_F_S2); FIELD(x2, FCP_F_S3); q.su_vol = x1; q.sp_vol = x2; }
  else {
    FIELD(te, FCP_F_TE); FIELD(ed, FCP_F_ED); FIELD(ms, FCP_F_MAGSTRAIN); FIELD(dnw, FCP_F_DNW);
    q.te = te; q.ed = ed; q.magStrain = ms; q.dnw = dnw;
    FIELD(gen, FCP_F_GEN);
    q.gen = gen;
    if (is_k) {
      FIELD(u, FCP_F_U); FIELD(v, FCP_F_V); FIELD(w, FCP_F_W); FIELD(visw, FCP_F_VISW); FIELD(tau, FCP_F_TAU);
      q.u = u; q.v = v; q.w = w; q.visw = visw; q.tau = tau;
    }
    if (is_sst) {
      FIELD(fsst, FCP_F_FSST); FIELD(wd, FCP_F_WALLDIST); FIELD(gte, FCP_F_DTEDXI);
      q.fsst = fsst; q.walldist = wd; q.gte = gte;
    }
  }
  if (prm->grad_method != FCP_GRAD_GAUSS && !ctx->Dmat[prm->grad_method]) FCP_TRY(fcp_create_lsq_grad_matrix(ctx, prm->grad_method));
  FCP_TRY(fcp_grad_opt(ctx, prm->grad_method, prm->limiter, phi_field, gfield));             // call grad(te, dTedxi)
  if (prm->kind == FCP_SC_OMEGA_SST) {                                                       // F1, k_omega_SST.f90:242-268
    FIELD(fsst, FCP_F_FSST);
    FCP_TRY(fvm_sst_blend(ctx, prm->viscos, q.walldist, q.gte, g, den, q.te, q.ed, fsst));
  }
  FCP_CUDA(cudaMemsetAsync(a, 0, sizeof(double) * (size_t)ctx->pat.nnzp, ctx->stream));      // a = 0
  if (ctx->comm) FCP_TRY(comm_exchange(ctx, vis, 1));       // ghost values of what facefluxsc reads across a process face (phi and its gradient: done by grad)
  if (ctx->comm && is_sst) { FIELD(fsst, FCP_F_FSST); FCP_TRY(comm_exchange(ctx, fsst, 1)); }   // ... and F1 of the cell across, when that cell owns the face
  FCP_TRY(fvm_sc_assemble(ctx, q));
  FCP_TRY(fcp_csrsolve(ctx, prm->solver, phi_field, FCP_F_SU, prm->maxiter, prm->tol_abs, prm->tol_rel, rep));
  FCP_TRY(fvm_update_boundary(ctx, phi));
  double *d_mm = nullptr, mm[2] = {0.0, 0.0};
  const int32_t cnt = is_sst ? ctx->nT : ctx->n;           // k_omega_SST.f90:776-786 takes the extrema of, and clips, the whole array
  FCP_TRY(fvm_minmax(ctx, phi, &d_mm, cnt));
  FCP_CUDA(cudaMemcpyAsync(mm, d_mm, sizeof(mm), cudaMemcpyDeviceToHost, ctx->stream));
  FCP_CUDA(cudaStreamSynchronize(ctx->stream));
  if (fimin) *fimin = mm[0];
  if (fimax) *fimax = mm[1];
  if (prm->kind != FCP_SC_GENERIC && mm[0] < 0.0) FCP_TRY(fvm_clip_small(ctx, phi, cnt));    // :430
  return FCP_OK;
}
// wall_distance   src/mesh/wall_distance.f90:75-133: laplacian(1, phi) with every patch Dirichlet 0, q = -vol, iccg(500, 1e-12, 1e-10), owner values into
// the non-wall boundary slots, grad_gauss, d = -|grad phi| + sqrt(|grad phi|^2 + 2 phi).  Uses S0 (phi), S1 (mu), SU (q), G0, A.
extern "C" int fcp_wall_distance(fcp_ctx *ctx, fcp_report *rep) {
  if (!ctx) return FCP_EINVAL;
  FCP_CUDA(cudaSetDevice(ctx->device));
  FCP_TRY(fcp_field_fill(ctx, FCP_F_S1, 1.0));                                            // :89  mu = 1
  FCP_TRY(fcp_field_fill(ctx, FCP_F_S0, 0.0));                                            // :86  phi = 0
  FCP_TRY(fcp_laplacian(ctx, FCP_F_S1, FCP_F_S0));                                        // :92
  FIELD(q, FCP_F_SU); FIELD(phi, FCP_F_S0); FIELD(g, FCP_F_G0); FIELD(wd, FCP_F_WALLDIST);
  FCP_TRY(fvm_neg_vol(ctx, q));                                                           // :83
  FCP_TRY(fcp_csrsolve(ctx, FCP_SOLVER_ICCG, FCP_F_S0, FCP_F_SU, 500, 1e-12, 1e-10, rep)); // :95-101
  FCP_TRY(fvm_wall_distance_finish(ctx, 0, phi, g, wd));                                  // :105-118
  FCP_TRY(fcp_grad(ctx, FCP_GRAD_GAUSS, FCP_F_S0, FCP_F_G0, 1));                          // :121
  return fvm_wall_distance_finish(ctx, 1, phi, g, wd);                                    // :124-125
}
extern "C" int fcp_calc_strain_and_vorticity(fcp_ctx *ctx) {
  if (!ctx) return FCP_EINVAL;
  FCP_CUDA(cudaSetDevice(ctx->device));
  FIELD(gu, FCP_F_DUDXI); FIELD(gv, FCP_F_DVDXI); FIELD(gw, FCP_F_DWDXI); FIELD(ms, FCP_F_MAGSTRAIN); FIELD(vo, FCP_F_VORTICITY);
  return fvm_strain(ctx, gu, gv, gw, ms, vo);
}
extern "C" int fcp_modify_mu_eff_k_epsilon_rlzb(fcp_ctx *ctx, double urfVis, double viscos) {
  if (!ctx) return FCP_EINVAL;
  FCP_CUDA(cudaSetDevice(ctx->device));      // cell- and boundary-face-local: nothing to exchange (the next consumer of vis exchanges its ghost values)
  FIELD(gu, FCP_F_DUDXI); FIELD(gv, FCP_F_DVDXI); FIELD(gw, FCP_F_DWDXI); FIELD(te, FCP_F_TE); FIELD(ed, FCP_F_ED); FIELD(den, FCP_F_DEN);
  FIELD(u, FCP_F_U); FIELD(v, FCP_F_V); FIELD(w, FCP_F_W); FIELD(dnw, FCP_F_DNW); FIELD(vis, FCP_F_VIS); FIELD(visw, FCP_F_VISW);
  FIELD(ypl, FCP_F_YPL); FIELD(tau, FCP_F_TAU);
  return fvm_mu_eff_rlzb(ctx, urfVis, viscos, gu, gv, gw, te, ed, den, u, v, w, dnw, vis, visw, ypl, tau);
}

extern "C" int fcp_modify_mu_eff_k_omega_sst(fcp_ctx *ctx, double urfVis, double viscos, double densit, int lowre) {
  if (!ctx) return FCP_EINVAL;
  FCP_CUDA(cudaSetDevice(ctx->device));
  FIELD(ms, FCP_F_MAGSTRAIN); FIELD(wd, FCP_F_WALLDIST); FIELD(te, FCP_F_TE); FIELD(ed, FCP_F_ED); FIELD(den, FCP_F_DEN);
  FIELD(u, FCP_F_U); FIELD(v, FCP_F_V); FIELD(w, FCP_F_W); FIELD(dnw, FCP_F_DNW); FIELD(vis, FCP_F_VIS); FIELD(visw, FCP_F_VISW);
  FIELD(ypl, FCP_F_YPL); FIELD(tau, FCP_F_TAU);
  return fvm_mu_eff_sst(ctx, urfVis, viscos, densit, lowre, ms, wd, te, ed, den, u, v, w, dnw, vis, visw, ypl, tau);
}
extern "C" int fcp_grad_gauss_fvx(fcp_ctx *ctx, int phi_field, int grad_field) {
  if (!ctx) return FCP_EINVAL;
  if (!fcp_is_gradient_field(grad_field) || phi_field < 0 || phi_field >= FCP_F_COUNT || (phi_field >= FCP_F_DUDXI && phi_field <= FCP_F_H) || fcp_is_gradient_field(phi_field)) {
    fcp_set_error("fcp_grad_gauss_fvx: bad field id");
    return FCP_EINVAL;
  }
  FCP_CUDA(cudaSetDevice(ctx->device));
  FIELD(phi, phi_field); FIELD(g, grad_field);
  FIELD(gtmp, grad_field == FCP_F_G1 ? FCP_F_G0 : FCP_F_G1);     // the first pass's gradient
  if (ctx->comm) FCP_TRY(comm_exchange(ctx, phi, 1));
  return fvm_grad_gauss_fvx(ctx, phi, gtmp, g);
}
// grad_gauss of the MPI tree, src-par/gradients.f90:1547-1664: nigrad passes of gradco (switchable where the two trees differ)
extern "C" int fcp_grad_gauss_iter(fcp_ctx *ctx, int phi_field, int grad_field, int nigrad) {
  if (!ctx) return FCP_EINVAL;
  if (!fcp_is_gradient_field(grad_field) || phi_field < 0 || phi_field >= FCP_F_COUNT || (phi_field >= FCP_F_DUDXI && phi_field <= FCP_F_H) || fcp_is_gradient_field(phi_field)) {
    fcp_set_error("fcp_grad_gauss_iter: bad field id");
    return FCP_EINVAL;
  }
  if (nigrad < 1) { fcp_set_error("fcp_grad_gauss_iter: nigrad must be >= 1 (the reference's DO loop would leave the gradient untouched)"); return FCP_EINVAL; }
  FCP_CUDA(cudaSetDevice(ctx->device));
  FIELD(phi, phi_field); FIELD(g, grad_field);
  FIELD(gtmp, grad_field == FCP_F_G1 ? FCP_F_G0 : FCP_F_G1);
  if (ctx->comm) FCP_TRY(comm_exchange(ctx, phi, 1));
  return fvm_grad_gauss_passes(ctx, phi, gtmp, g, nigrad);
}
extern "C" int fcp_modify_viscosity_sgs(fcp_ctx *ctx, int model, double urfVis, double viscos) {
  if (!ctx) return FCP_EINVAL;
  if (model != FCP_SGS_WALE && model != FCP_SGS_VREMAN) { fcp_set_error("fcp_modify_viscosity_sgs: unknown model %d", model); return FCP_EINVAL; }
  FCP_CUDA(cudaSetDevice(ctx->device));
  FCP_TRY(fcp_grad_gauss_fvx(ctx, FCP_F_U, FCP_F_DUDXI));
  FCP_TRY(fcp_grad_gauss_fvx(ctx, FCP_F_V, FCP_F_DVDXI));
  FCP_TRY(fcp_grad_gauss_fvx(ctx, FCP_F_W, FCP_F_DWDXI));
  FIELD(gu, FCP_F_DUDXI); FIELD(gv, FCP_F_DVDXI); FIELD(gw, FCP_F_DWDXI); FIELD(den, FCP_F_DEN); FIELD(vis, FCP_F_VIS); FIELD(visw, FCP_F_VISW);
  return fvm_sgs_viscosity(ctx, model, urfVis, viscos, gu, gv, gw, den, vis, visw);
}

// calcp_piso   Pressure/calcp_piso.f90:81-489
extern "C" int fcp_calcp_piso(fcp_ctx *ctx, const fcp_piso_params *prm, fcp_report *rep) {
  if (!ctx || !prm) return FCP_EINVAL;
  if (prm->ncorr < 1 || prm->npcor < 1) { fcp_set_error("calcp_piso: ncorr and npcor must be >= 1"); return FCP_EINVAL; }
  if (prm->pscheme < 0 || prm->pscheme > 2) { fcp_set_error("unknown pscheme %d", prm->pscheme); return FCP_EINVAL; }
  FCP_CUDA(cudaSetDevice(ctx->device));
  FIELD(den, FCP_F_DEN); FIELD(u, FCP_F_U); FIELD(v, FCP_F_V); FIELD(w, FCP_F_W); FIELD(p, FCP_F_P); FIELD(pp, FCP_F_PP);
  FIELD(g, FCP_F_DPDXI); FIELD(apu, FCP_F_APU); FIELD(apv, FCP_F_APV); FIELD(apw, FCP_F_APW); FIELD(a, FCP_F_A); FIELD(h, FCP_F_H);
  FIELD(su, FCP_F_SU); FIELD(sv, FCP_F_SV); FIELD(sw, FCP_F_SW); FIELD(fl, FCP_F_FLMASS);
  FIELD(rU, FCP_F_RU); FIELD(rV, FCP_F_RV); FIELD(rW, FCP_F_RW);
  if (!ctx->d_sum) FCP_TRY(dev_alloc(&ctx->d_sum, 8));
  double ncells = (double)ctx->n;
  if (ctx->comm) FCP_TRY(fcp_global_sum(ctx, &ncells));
  FCP_CUDA(cudaMemcpyAsync(h, a, sizeof(double) * (size_t)ctx->pat.nnzp, cudaMemcpyDeviceToDevice, ctx->stream));   // :81  h = a
  for (int icorr = 1; icorr <= prm->ncorr; ++icorr) {
    if (ctx->comm) { FCP_TRY(comm_exchange(ctx, u, 1)); FCP_TRY(comm_exchange(ctx, v, 1)); FCP_TRY(comm_exchange(ctx, w, 1)); }
    FCP_TRY(fvm_piso_hbya(ctx, h, rU, rV, rW, apu, apv, apw, u, v, w, su, sv, sw));                                 // :102-129
    if (!prm->const_mflux && ctx->g_outlet) {                                                                      // :186 (the outlet faces do not read a or su: order is free)
      FCP_TRY(ensure_outlet_list(ctx));
      FCP_TRY(fvm_adjust_mass_flow(ctx, ctx->nout, ctx->d_oface, den, u, v, w, fl, prm->flomas));
    }
    if (ctx->comm) {
      double *sc[] = {den, u, v, w, p, apu};
      for (double *x : sc) FCP_TRY(comm_exchange(ctx, x, 1));
      FCP_TRY(comm_exchange(ctx, g, 3));
    }
    AsmArgs args{den, u, v, w, p, g, apu, apv, apw, pp, u, v, w, a, su, fl};
    FCP_TRY(fvm_assemble_pcorr(ctx, args, true));                                                                    // :140-297
    for (int ipcorr = 1; ipcorr <= prm->npcor; ++ipcorr) {                                                           // :308
      FCP_TRY(fcp_csrsolve(ctx, prm->solver, FCP_F_PP, FCP_F_SU, prm->maxiter, prm->tol_abs, prm->tol_rel,
                           rep ? &rep[(icorr - 1) * prm->npcor + (ipcorr - 1)] : nullptr));
      FCP_TRY(fvm_sum(ctx, pp, ctx->d_sum));                                                                         // :330
      if (ctx->comm) FCP_TRY(comm_allgather_sum(ctx->comm, ctx->d_sum, 1, ctx->stream));
      FCP_TRY(fvm_piso_pupdate(ctx, ncells, prm->urfp, ctx->d_sum, pp, p));                                          // :333
      if (ipcorr != prm->npcor) {
        // :336-344 bpres + grad_gauss twice == the fused two-stage kernel with linear interpolation and no sources
        if (ctx->comm) FCP_TRY(comm_exchange(ctx, p, 1));
        FCP_TRY(fvm_gradp(ctx, FCP_PSCHEME_LINEAR, p, apu, nullptr, nullptr, nullptr, g, nullptr, nullptr));
        if (ctx->comm) { FCP_TRY(comm_exchange(ctx, g, 3)); }
        FCP_TRY(fvm_piso_fluxmc(ctx, den, apu, g, su));                                                              // :349-364
      } else {
        // last pass: the gradient of :336-344 is recomputed bit-identically by gradp_and_sources below (p unchanged)
        if (ctx->comm) FCP_TRY(comm_exchange(ctx, p, 1));
        FCP_TRY(fvm_correct_flux(ctx, a, p, fl));                                                                    // :377-387
        if (ctx->npro) FCP_TRY(fvm_correct_flux_proc(ctx, a, p, fl));
      }
    }
    CorrectArgs ca{u, v, w, nullptr, apu, apv, apw, 0.0, nullptr};
    FCP_TRY(gradp_impl(ctx, prm->pscheme, p, &ca));                                                                  // :425-431, :485
    if (ctx->nper) FCP_TRY(fvm_correct_flux_periodic(ctx, a, p, fl));                                                // :441-460 (the whole pressure)
    if (ctx->has_pressure_patch) FCP_TRY(fvm_correct_pressure_bnd(ctx, den, apu, pp, u, v, w, fl));                  // :466-479
  }
  return FCP_OK;
}

// ---------------------------------------------------------------------------------------------
// explicit-CSR solver objects
// ---------------------------------------------------------------------------------------------
extern "C" int fcp_solver_create(int32_t n, int32_t nnz, const int32_t *ia, const int32_t *ja, const int32_t *diag, int device, fcp_solver **out) {
  if (!ia || !ja || !diag || !out || n < 0) { fcp_set_error("fcp_solver_create: bad argument"); return FCP_EINVAL; }
  *out = nullptr;
  if (ia[0] != 1 || ia[n] - 1 != nnz) { fcp_set_error("fcp_solver_create: ia must be 1-based with ia(n+1)-1 == nnz"); return FCP_EINVAL; }
  for (int32_t i = 0; i < n; ++i) {
    if (diag[i] < ia[i] || diag[i] >= ia[i + 1] || ja[diag[i] - 1] != i + 1) { fcp_set_error("row %d: diag does not point at a(i,i)", i + 1); return FCP_EINVAL; }
    for (int32_t k = ia[i]; k < ia[i + 1]; ++k) {
      if (ja[k - 1] < 1 || ja[k - 1] > n) { fcp_set_error("row %d: column out of range", i + 1); return FCP_EINVAL; }
      if (k > ia[i] && ja[k - 1] <= ja[k - 2]) { fcp_set_error("row %d: columns must be ascending (create_CSR_matrix order)", i + 1); return FCP_EINVAL; }
    }
  }
  FCP_TRY(select_device(device));
  fcp_solver *s = new fcp_solver();
  s->device = device;
  FCP_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  int rc = sell_from_csr(s->pat, n, n, ia, ja, diag, nullptr);
  if (rc != FCP_OK) { delete s; return rc; }
  FCP_TRY(dev_alloc(&s->a, (size_t)s->pat.nnzp));
  FCP_CUDA(cudaMemset(s->a, 0, sizeof(double) * (size_t)std::max<int64_t>(s->pat.nnzp, 1)));
  FCP_CUDA(cudaStreamSynchronize(0));
  FCP_TRY(dev_alloc(&s->a_csr, (size_t)nnz));
  FCP_TRY(dev_alloc(&s->fi, (size_t)n));
  FCP_TRY(dev_alloc(&s->rhs, (size_t)n));
  *out = s;
  return FCP_OK;
}
extern "C" int fcp_solver_destroy(fcp_solver *s) {
  if (!s) return FCP_OK;
  cudaSetDevice(s->device);
  cudaFree(s->a); cudaFree(s->a_csr); cudaFree(s->fi); cudaFree(s->rhs);
  sell_free(s->pat);
  krylov_ws_free(s->ws);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return FCP_OK;
}
extern "C" int fcp_solver_solve(fcp_solver *s, int solver, const double *a, double *fi, const double *rhs, int32_t itr_max, double tol_abs,
                                double tol_rel, fcp_report *rep) {
  if (!s || !a || !fi || !rhs) return FCP_EINVAL;
  FCP_CUDA(cudaSetDevice(s->device));
  const size_t n = s->pat.n;
  FCP_CUDA(cudaMemcpyAsync(s->a_csr, a, sizeof(double) * (size_t)s->pat.nnz, cudaMemcpyHostToDevice, s->stream));
  FCP_CUDA(cudaMemcpyAsync(s->fi, fi, sizeof(double) * n, cudaMemcpyHostToDevice, s->stream));
  FCP_CUDA(cudaMemcpyAsync(s->rhs, rhs, sizeof(double) * n, cudaMemcpyHostToDevice, s->stream));
  FCP_TRY(sell_values_from_csr(s->pat, s->a_csr, s->a, s->stream));
  FCP_TRY(krylov_solve(solver, s->pat, s->a, s->fi, s->rhs, s->ws, itr_max, tol_abs, tol_rel, rep, s->stream, nullptr, nullptr));
  FCP_CUDA(cudaMemcpyAsync(fi, s->fi, sizeof(double) * n, cudaMemcpyDeviceToHost, s->stream));
  FCP_CUDA(cudaStreamSynchronize(s->stream));
  return FCP_OK;
}

extern "C" int fcp_profile_enable(fcp_ctx *ctx, int on) {
  if (!ctx) return FCP_EINVAL;
  ctx->prof.resolve();
  ctx->prof.on = on != 0;
  return FCP_OK;
}
extern "C" int fcp_profile_reset(fcp_ctx *ctx) {
  if (!ctx) return FCP_EINVAL;
  ctx->prof.reset();
  return FCP_OK;
}
extern "C" int fcp_profile_read(fcp_ctx *ctx, int kclass, double *total_ms, int64_t *launches) {
  if (!ctx || kclass < 0 || kclass >= FCP_K_COUNT) return FCP_EINVAL;
  ctx->prof.resolve();
  if (total_ms) *total_ms = ctx->prof.total_ms[kclass];
  if (launches) *launches = ctx->prof.launches[kclass];
  return FCP_OK;
}

// ---------------------------------------------------------------------------------------------
// timing helpers
// ---------------------------------------------------------------------------------------------
extern "C" int fcp_timer_start(fcp_ctx *ctx) {
  if (!ctx) return FCP_EINVAL;
  FCP_CUDA(cudaEventRecord(ctx->t0, ctx->stream));
  return FCP_OK;
}
extern "C" int fcp_timer_stop(fcp_ctx *ctx, float *ms) {
  if (!ctx || !ms) return FCP_EINVAL;
  FCP_CUDA(cudaEventRecord(ctx->t1, ctx->stream));
  FCP_CUDA(cudaEventSynchronize(ctx->t1));
  FCP_CUDA(cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
  return FCP_OK;
}
extern "C" int fcp_flush_l2(fcp_ctx *ctx) {
  if (!ctx) return FCP_EINVAL;
  const size_t bytes = (size_t)256 << 20;
  if (!ctx->flushbuf) FCP_CUDA(cudaMalloc((void **)&ctx->flushbuf, bytes));
  k_fill<<<148 * 8, 256, 0, ctx->stream>>>((int64_t)(bytes / sizeof(double)), ctx->flushbuf, 1.0);
  FCP_CHECK_LAUNCH();
  return FCP_OK;
}
