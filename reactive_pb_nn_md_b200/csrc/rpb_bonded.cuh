// Per-molecule bonded + intramolecular non-bonded terms, evaluated by one thread on a molecule
// image held in registers/local memory (molecules have <= RPB_MA atoms).
//   bonds / angles / dihedrals   intra_bonded_interactions.f90:84-552
//   intramolecular pairs         pair_int_real_space.f90:386-588 (+ intra_pme_exclusion :781-816)
#pragma once
#include "rpb_dev.cuh"

__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

struct MolEnergies { double e_bond, e_angle, e_dih, e_elec, e_vdw; };

// number of independent terms of a molecule: bonds, angles, dihedrals, intramolecular atom pairs (in that order)
__device__ __forceinline__ int molecule_term_count(const MolTypeDev& T, int n_atom, bool bonded, bool nonbonded) {
  return (bonded ? T.n_bond + T.n_angle + T.n_dih : 0) + ((nonbonded && n_atom > 1) ? n_atom * (n_atom - 1) / 2 : 0);
}

// ONE term (index as counted by molecule_term_count): x[a][3], f[a][3] and E are accumulated into, type[a] actual atom
// types, q[a] charges.  molecule_terms() runs them in order on one thread; the MS-EVB item kernel gives a term to a lane.
__device__ inline void molecule_term(const Dev& d, const MolTypeDev& T, int n_atom, const double (*x)[3], const int* type,
                                     const double* q, double (*f)[3], MolEnergies& E, bool bonded, bool nonbonded, int term) {
  if (bonded) {
    if (term < T.n_bond) {
      const int b = term;
      int i = T.bond[b][0], j = T.bond[b][1];
      double r[3] = {x[i][0] - x[j][0], x[i][1] - x[j][1], x[i][2] - x[j][2]};
      double rm = sqrt(dot3(r, r));
      const double* P = T.bond_par[b];
      double e, s;  // force_ij = s * r
      if (T.bond_kind[b] == 1) {
        e = 0.5 * P[1] * ((rm - P[0]) * (rm - P[0]));
        s = -P[1] * (rm - P[0]) / rm;
      } else if (T.bond_kind[b] == 2) {
        double t = rm * rm - P[0] * P[0];
        e = 0.25 * P[1] * (t * t);
        s = -P[1] * t;
      } else {
        double ex = exp(-P[1] * (rm - P[2]));
        e = P[0] * ((1.0 - ex) * (1.0 - ex));
        s = -2.0 * P[0] * P[1] * ex * (1.0 - ex) / rm;
      }
      E.e_bond += e;
      for (int k = 0; k < 3; k++) { f[i][k] += s * r[k]; f[j][k] -= s * r[k]; }
      return;
    }
    term -= T.n_bond;
    if (term < T.n_angle) {
      const int a = term;
      int i = T.angle[a][0], j = T.angle[a][1], k = T.angle[a][2];
      double rij[3], rkj[3];
      for (int c = 0; c < 3; c++) { rij[c] = x[i][c] - x[j][c]; rkj[c] = x[k][c] - x[j][c]; }
      double rijm = sqrt(dot3(rij, rij)), rkjm = sqrt(dot3(rkj, rkj));
      double cosine = dot3(rij, rkj) / rijm / rkjm;
      double th0 = T.angle_par[a][0], cth = T.angle_par[a][1], fac, e;
      if (T.angle_kind[a] == 1) {
        double theta;
        if (cosine < -0.999999999) theta = d.pi; else if (cosine > 0.999999999) theta = 0.0; else theta = acos(cosine);
        e = 0.5 * cth * ((theta - th0) * (theta - th0));
        fac = (fabs(theta - th0) < 1e-4) ? 0.0 : cth * (theta - th0) / sqrt(1.0 - cosine * cosine);
      } else {
        double c0 = cos(th0);
        e = 0.5 * cth * ((cosine - c0) * (cosine - c0));
        fac = -cth * (cosine - c0);
      }
      E.e_angle += e;
      for (int c = 0; c < 3; c++) {
        double fij = fac * (rkj[c] / rijm / rkjm - cosine * rij[c] / (rijm * rijm));
        double fkj = fac * (rij[c] / rijm / rkjm - cosine * rkj[c] / (rkjm * rkjm));
        f[i][c] += fij; f[k][c] += fkj; f[j][c] = f[j][c] - fij - fkj;
      }
      return;
    }
    term -= T.n_angle;
    if (term < T.n_dih) {
      const int q4 = term;
      int i = T.dih[q4][0], j = T.dih[q4][1], k = T.dih[q4][2], l = T.dih[q4][3];
      double rji[3], rkj[3], rlk[3];
      for (int c = 0; c < 3; c++) { rji[c] = x[j][c] - x[i][c]; rkj[c] = x[k][c] - x[j][c]; rlk[c] = x[l][c] - x[k][c]; }
      double rji2 = dot3(rji, rji), rkj2 = dot3(rkj, rkj), rlk2 = dot3(rlk, rlk);
      double dkj_ji = dot3(rkj, rji), dlk_kj = dot3(rlk, rkj), dlk_ji = dot3(rlk, rji);
      double ab = dkj_ji * dlk_kj - dlk_ji * rkj2;
      double aa = rji2 * rkj2 - dkj_ji * dkj_ji;
      double bb = rlk2 * rkj2 - dlk_kj * dlk_kj;
      double sa = sqrt(aa), sb = sqrt(bb);
      double cosine = ab / sa / sb;
      double xi;
      if (cosine < -0.999999999) xi = d.pi; else if (cosine > 0.999999999) xi = 0.0; else xi = acos(cosine);
      const double* P = T.dih_par[q4];
      double fac = 0.0, e = 0.0;
      int shift = 0, kind = T.dih_kind[q4];
      if (kind == 1) {
        e = P[1] * (1.0 + cos(P[2] * xi - P[0]));
        double c2 = cosine * cosine;
        if (fabs(c2 - 1.0) < 1e-4) fac = 0.0;  // reference stops if the phase is not 0/pi (:436-441)
        else fac = P[1] * -sin(P[2] * xi - P[0]) * P[2] / sqrt(1.0 - c2);
      } else if (kind == 2) {
        if (xi > (d.pi / 2.0)) { xi = fabs(xi - d.pi); shift = 1; }
        e = 0.5 * P[1] * ((xi - P[0]) * (xi - P[0]));
        fac = (fabs(xi - P[0]) < 1e-4) ? 0.0 : P[1] * (xi - P[0]) / sqrt(1.0 - cosine * cosine);
      } else if (kind == 3) {
        double c2 = cosine * cosine, c3 = c2 * cosine, c4 = c3 * cosine, c5 = c4 * cosine;
        e = P[0] - P[1] * cosine + P[2] * c2 - P[3] * c3 + P[4] * c4 - P[5] * c5;
        fac = P[1] - 2.0 * P[2] * cosine + 3.0 * P[3] * c2 - 4.0 * P[4] * c3 + 5.0 * P[5] * c4;
      } else {
        return;
      }
      E.e_dih += e;
      double aa15 = pow(aa, 1.5), bb15 = pow(bb, 1.5);
      double sgn = shift ? -1.0 : 1.0;
      for (int c = 0; c < 3; c++) {
        double dab_ji = rkj[c] * dlk_kj - rlk[c] * rkj2;
        double dab_kj = rji[c] * dlk_kj + dkj_ji * rlk[c] - dlk_ji * 2.0 * rkj[c];
        double dab_lk = dkj_ji * rkj[c] - rji[c] * rkj2;
        double daa_ji = rkj2 * 2.0 * rji[c] - 2.0 * dkj_ji * rkj[c];
        double daa_kj = rji2 * 2.0 * rkj[c] - 2.0 * dkj_ji * rji[c];
        double dbb_kj = rlk2 * 2.0 * rkj[c] - 2.0 * dlk_kj * rlk[c];
        double dbb_lk = rkj2 * 2.0 * rlk[c] - 2.0 * dlk_kj * rkj[c];
        double fji = sgn * fac * (dab_ji / sa / sb - 0.5 * ab / aa15 / sb * daa_ji - 0.5 * ab / sa / bb15 * 0.0);
        double fkj = sgn * fac * (dab_kj / sa / sb - 0.5 * ab / aa15 / sb * daa_kj - 0.5 * ab / sa / bb15 * dbb_kj);
        double flk = sgn * fac * (dab_lk / sa / sb - 0.5 * ab / aa15 / sb * 0.0 - 0.5 * ab / sa / bb15 * dbb_lk);
        f[i][c] -= fji; f[j][c] = f[j][c] + fji - fkj; f[k][c] = f[k][c] + fkj - flk; f[l][c] += flk;
      }
      return;
    }
    term -= T.n_dih;
  }
  if (nonbonded && n_atom > 1) {
    {
      {
        int i = 0;
        while (term >= n_atom - 1 - i) { term -= n_atom - 1 - i; i++; }     // pairs (i, j > i), row by row
        const int j = i + 1 + term;
        double dr[3] = {x[i][0] - x[j][0], x[i][1] - x[j][1], x[i][2] - x[j][2]};
        double dr2 = dot3(dr, dr);
        double qq = q[i] * q[j];
        double e_el = 0.0, e_vdw = 0.0, ff[3];
        int ex = T.pair_excl[i][j];
        if (ex == 1) {
          excl_terms(d, dr, dr2, qq, e_el, ff);
        } else {
          int pidx = type[i] * d.nT + type[j];
          int vt = d.vdw_type[pidx];
          const double* par = (ex == 2 && vt == 0) ? &d.vdw_param14[6 * pidx] : &d.vdw_param[6 * pidx];
          pair_terms(d, dr, dr2, qq, vt, par, dr2 < d.rc2, e_el, e_vdw, ff);
        }
        E.e_elec += e_el; E.e_vdw += e_vdw;
        for (int k = 0; k < 3; k++) { f[i][k] += ff[k]; f[j][k] -= ff[k]; }
      }
    }
  }
}

__device__ inline void molecule_terms(const Dev& d, const MolTypeDev& T, int n_atom, const double (*x)[3], const int* type,
                                      const double* q, double (*f)[3], MolEnergies& E, bool bonded, bool nonbonded) {
  E.e_bond = E.e_angle = E.e_dih = E.e_elec = E.e_vdw = 0.0;
  const int n = molecule_term_count(T, n_atom, bonded, nonbonded);
  for (int t = 0; t < n; t++) molecule_term(d, T, n_atom, x, type, q, f, E, bonded, nonbonded, t);
}
