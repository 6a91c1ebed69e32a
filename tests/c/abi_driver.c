/* Plain-C driver of the rpbmd C-ABI (no ctypes, no Python in the loop): the call order of the Fortran shim
 * (fortran/rpbmd_iso_c.f90: rpb_setup -> rpb_push_state -> rpb_force_energy -> rpb_pull_results), on a 2-ion system
 * whose answer is known in closed form, plus compile-time checks that rpb_config has the layout the shim's
 * `type, bind(C) :: rpb_config` implies (c_int = 4 bytes, c_double = 8 bytes, natural alignment, declaration order).
 *
 *   gcc -std=c11 -I include tests/c/abi_driver.c -o abi_driver -ldl -lm && ./abi_driver <library.so>
 * Exit status 0 = every check passed.  Linked against nothing: the library is dlopen'ed, like a Fortran driver's
 * shared-library dependency would be resolved at run time. */
#include <assert.h>
#include <dlfcn.h>
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "rpbmd.h"

/* ---- layout of rpb_config as the bind(C) derived type lays it out ---- */
#define INT_FIELD(k) (4 * (k))
_Static_assert(sizeof(int) == 4 && sizeof(double) == 8, "c_int / c_double");
_Static_assert(offsetof(rpb_config, n_atoms) == INT_FIELD(0), "n_atoms");
_Static_assert(offsetof(rpb_config, n_mole) == INT_FIELD(1), "n_mole");
_Static_assert(offsetof(rpb_config, n_atom_type) == INT_FIELD(2), "n_atom_type");
_Static_assert(offsetof(rpb_config, n_mole_type) == INT_FIELD(3), "n_mole_type");
_Static_assert(offsetof(rpb_config, pme_grid) == INT_FIELD(4), "pme_grid");
_Static_assert(offsetof(rpb_config, spline_order) == INT_FIELD(5), "spline_order");
_Static_assert(offsetof(rpb_config, spline_grid) == INT_FIELD(6), "spline_grid");
_Static_assert(offsetof(rpb_config, erfc_grid) == INT_FIELD(7), "erfc_grid");
_Static_assert(offsetof(rpb_config, tt_grid) == INT_FIELD(8), "tt_grid");
_Static_assert(offsetof(rpb_config, na_nslist) == INT_FIELD(9), "na_nslist");
_Static_assert(offsetof(rpb_config, nb_nslist) == INT_FIELD(10), "nb_nslist");
_Static_assert(offsetof(rpb_config, nc_nslist) == INT_FIELD(11), "nc_nslist");
_Static_assert(offsetof(rpb_config, verlet_capacity) == INT_FIELD(12), "verlet_capacity");
_Static_assert(offsetof(rpb_config, device) == INT_FIELD(13), "device");
_Static_assert(offsetof(rpb_config, rank) == INT_FIELD(14), "rank");
_Static_assert(offsetof(rpb_config, world_size) == INT_FIELD(15), "world_size");
_Static_assert(offsetof(rpb_config, n_threads) == INT_FIELD(16), "n_threads");
_Static_assert(offsetof(rpb_config, evb_max_chain) == INT_FIELD(17), "evb_max_chain");
_Static_assert(offsetof(rpb_config, evb_max_states) == INT_FIELD(18), "evb_max_states");
_Static_assert(offsetof(rpb_config, reserved_i) == INT_FIELD(19), "reserved_i(4)");
/* 23 ints = 92 bytes; the first double is aligned to 96 */
#define DBL_FIELD(k) (96 + 8 * (k))
_Static_assert(offsetof(rpb_config, box) == DBL_FIELD(0), "box(9)");
_Static_assert(offsetof(rpb_config, alpha_sqrt) == DBL_FIELD(9), "alpha_sqrt");
_Static_assert(offsetof(rpb_config, real_space_cutoff) == DBL_FIELD(10), "real_space_cutoff");
_Static_assert(offsetof(rpb_config, verlet_cutoff) == DBL_FIELD(11), "verlet_cutoff");
_Static_assert(offsetof(rpb_config, delta_t) == DBL_FIELD(12), "delta_t");
_Static_assert(offsetof(rpb_config, erfc_dx) == DBL_FIELD(13), "erfc_dx");
_Static_assert(offsetof(rpb_config, tt_max) == DBL_FIELD(14), "tt_max");
_Static_assert(offsetof(rpb_config, pi) == DBL_FIELD(15), "pi");
_Static_assert(offsetof(rpb_config, pi_sqrt) == DBL_FIELD(16), "pi_sqrt");
_Static_assert(offsetof(rpb_config, conv_e2A_kJmol) == DBL_FIELD(17), "conv_e2A_kJmol");
_Static_assert(offsetof(rpb_config, conv_kJmol_ang2ps2gmol) == DBL_FIELD(18), "conv_kJmol_ang2ps2gmol");
_Static_assert(offsetof(rpb_config, safe_verlet) == DBL_FIELD(19), "safe_verlet");
_Static_assert(offsetof(rpb_config, verlet_thresh) == DBL_FIELD(20), "verlet_thresh");
_Static_assert(offsetof(rpb_config, evb_first_solvation_cutoff) == DBL_FIELD(21), "evb_first_solvation_cutoff");
_Static_assert(offsetof(rpb_config, evb_reactive_pair_distance) == DBL_FIELD(22), "evb_reactive_pair_distance");
_Static_assert(offsetof(rpb_config, ewald_self) == DBL_FIELD(23), "ewald_self");
_Static_assert(offsetof(rpb_config, reserved_d) == DBL_FIELD(24), "reserved_d(4)");
_Static_assert(sizeof(rpb_config) == DBL_FIELD(28), "sizeof(rpb_config)");
_Static_assert(sizeof(rpb_energies) == 8 * 8, "sizeof(rpb_energies)");

#define T RPB_MAX_N_ATOM_TYPE
#define MT RPB_MAX_N_MOLE_TYPE
#define MA RPB_MAX_MOLE_ATOMS

static void* lib;
#define SYM(name) __typeof__(&name) p_##name = (__typeof__(&name))dlsym(lib, #name); if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }
#define CHECK(call) do { int rc__ = (call); if (rc__ != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc__, p_rpb_last_error(ctx)); return 3; } } while (0)

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: abi_driver <librpbmd(.so|_oracle.so)> [symbols-only]\n"); return 1; }
  lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  SYM(rpb_last_error) SYM(rpb_backend) SYM(rpb_create) SYM(rpb_destroy) SYM(rpb_set_tables) SYM(rpb_set_forcefield)
  SYM(rpb_set_molecule_types) SYM(rpb_set_evb) SYM(rpb_upload_state) SYM(rpb_initialize) SYM(rpb_force_energy) SYM(rpb_step)
  SYM(rpb_get_energies) SYM(rpb_download_state) SYM(rpb_get_r_com) SYM(rpb_get_neighbor_list) SYM(rpb_get_evb)
  SYM(rpb_peer_export) SYM(rpb_peer_import) SYM(rpb_peer_enabled)
  printf("backend: %s\n", p_rpb_backend());
  if (argc > 2) { printf("symbols ok\n"); return 0; }          /* no compute: the CUDA library on a box without a GPU */
  (void)p_rpb_set_evb; (void)p_rpb_get_evb; (void)p_rpb_peer_export; (void)p_rpb_peer_import; (void)p_rpb_peer_enabled; (void)p_rpb_get_neighbor_list;

  /* ---- rpb_setup: two point charges +1 / -1, 3 A apart, in a 32 A cubic box; no LJ, no bonded terms ---- */
  const int N = 2, M = 2, K = 32, G = 100000, TTG = 1000;
  const double L = 32.0, rc = 10.0, alpha = 0.3, conv = 1389.3546142578125, pi_sqrt = 1.772453851, pi = 3.141592654;
  rpb_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.n_atoms = N; cfg.n_mole = M; cfg.n_atom_type = 2; cfg.n_mole_type = 2; cfg.pme_grid = K; cfg.spline_order = 6;
  cfg.spline_grid = G; cfg.erfc_grid = G; cfg.tt_grid = TTG; cfg.na_nslist = cfg.nb_nslist = cfg.nc_nslist = 10;
  cfg.world_size = 1; cfg.n_threads = 1; cfg.evb_max_chain = RPB_EVB_MAX_CHAIN; cfg.evb_max_states = RPB_EVB_MAX_STATES;
  cfg.box[0] = cfg.box[4] = cfg.box[8] = L;
  cfg.alpha_sqrt = alpha; cfg.real_space_cutoff = rc; cfg.verlet_cutoff = 12.0; cfg.delta_t = 0.0005; cfg.erfc_dx = rc / G; cfg.tt_max = 50.0;
  cfg.pi = pi; cfg.pi_sqrt = pi_sqrt; cfg.conv_e2A_kJmol = conv; cfg.conv_kJmol_ang2ps2gmol = 100.0;
  cfg.safe_verlet = 1.2000000476837158; cfg.verlet_thresh = 1.2; cfg.evb_first_solvation_cutoff = 5.0; cfg.evb_reactive_pair_distance = 2.5;
  cfg.ewald_self = -2.0 * alpha / pi_sqrt * conv;                 /* update_Ewald_self pme.f90:608-625: -sum q^2 alpha/sqrt(pi) */
  rpb_ctx* ctx = NULL;
  CHECK(p_rpb_create(&ctx, &cfg));

  /* tables of initialize_energy_force (initialize_routines.f90:212-264); the B-splines analytically (this driver checks
   * the call sequence and the real-space / self terms, not the REAL*4 table quirks), CB = 0: no reciprocal part */
  double* B6 = calloc(G, 8); double* B5 = calloc(G, 8); double* et = calloc(G + 1, 8); double* st = calloc(G + 1, 8);
  double* tt = calloc(4 * TTG, 8); double* dtt = calloc(4 * TTG, 8); double* CB = calloc((size_t)K * K * K, 8);
  for (int i = 1; i <= G + 1; i++) { double r = i * cfg.erfc_dx, x = r * alpha; et[i - 1] = erfc(x) * conv; st[i - 1] = et[i - 1] + x * 2.0 / pi_sqrt * exp(-x * x) * conv; }
  for (int i = 0; i < G; i++) { B6[i] = 1.0 / 6.0; B5[i] = 1.0 / 5.0; }     /* any bounded weights: they multiply CB = 0 */
  CHECK(p_rpb_set_tables(ctx, B6, B5, et, st, tt, dtt, CB));

  double* vdw_p = calloc((size_t)T * T * 6, 8); double* vdw_p14 = calloc((size_t)T * T * 6, 8); int* vdw_t = calloc((size_t)T * T, 4);
  for (int i = 0; i < T * T; i++) vdw_t[i] = -1;                               /* no van der Waals term between any pair */
  double chg[T] = {1.0, -1.0}; int freeze[T] = {0};
  int* bt = calloc((size_t)T * T, 4); double* bp = calloc((size_t)T * T * 3, 8); int* at = calloc((size_t)T * T * T, 4); double* ap = calloc((size_t)T * T * T * 2, 8);
  int* dt = calloc((size_t)T * T * T * T, 4); double* dp = calloc((size_t)T * T * T * T * 6, 8);
  CHECK(p_rpb_set_forcefield(ctx, vdw_p, vdw_t, vdw_p14, chg, freeze, bt, bp, at, ap, dt, dp));
  int mt_natom[MT] = {1, 1}, mt_atype[MT * MA] = {0}, mt_n0[MT] = {0}, dummy[4] = {0}, excl[MT * MA * MA] = {0}, rp[MT * MA] = {0}, rb[MT * MA] = {0};
  mt_atype[0] = 1; mt_atype[MA] = 2;
  CHECK(p_rpb_set_molecule_types(ctx, mt_natom, mt_atype, mt_n0, dummy, mt_n0, dummy, mt_n0, dummy, excl, rp, rb));

  /* ---- rpb_push_state ---- */
  const double r = 3.0;
  double xyz[6] = {10.0, 10.0, 10.0, 10.0 + r, 10.0, 10.0}, vel[6] = {0}, mass[2] = {22.99, 35.45}, q[2] = {1.0, -1.0};
  int atype[2] = {1, 2}, mfirst[2] = {1, 2}, mnatom[2] = {1, 1}, mtype[2] = {1, 2};
  CHECK(p_rpb_upload_state(ctx, xyz, vel, mass, q, atype, mfirst, mnatom, mtype, 0));
  CHECK(p_rpb_initialize(ctx));

  /* ---- rpb_force_energy + rpb_pull_results ---- */
  CHECK(p_rpb_force_energy(ctx, 0));
  double force[6], xo[6], vo[6], mo[2], qo[2], rcom[6]; int ato[2], mfo[2], mno[2], mto[2], hyd = -1;
  CHECK(p_rpb_download_state(ctx, xo, vo, force, mo, qo, ato, mfo, mno, mto, &hyd));
  rpb_energies e;
  CHECK(p_rpb_get_energies(ctx, &e));
  CHECK(p_rpb_get_r_com(ctx, rcom));
  /* closed form with the reference's table semantics (effective argument r + dx): tolerance 1e-6 relative */
  const double e_real = -erfc(alpha * r) / r * conv, e_expect = e_real + cfg.ewald_self;
  const double f_expect = -(erfc(alpha * r) / (r * r) + 2.0 * alpha / pi_sqrt * exp(-alpha * alpha * r * r) / r) * conv;   /* on atom 2 along +x: attraction */
  printf("E_elec %.10f expected %.10f | F2x %.8f expected %.8f | E_recip %.3e\n", e.E_elec, e_expect, force[3], f_expect, e.E_recip);
  int ok = fabs(e.E_elec - e_expect) < 1e-4 * fabs(e_expect) && fabs(force[3] - f_expect) < 1e-4 * fabs(f_expect) &&
           fabs(force[0] + force[3]) < 1e-9 && fabs(e.E_recip) < 1e-12 && fabs(e.E_vdw) < 1e-12 &&
           ato[0] == 1 && ato[1] == 2 && mfo[0] == 1 && mfo[1] == 2 && mno[0] == 1 && mto[1] == 2 && hyd == 0 &&
           fabs(xo[3] - xyz[3]) < 1e-12 && fabs(rcom[3] - xyz[3]) < 1e-12 && qo[1] == -1.0 && mo[0] == 22.99;
  /* the same step through rpb_step (md_integrate_atomic): the ions attract */
  CHECK(p_rpb_step(ctx, 5, 0));
  CHECK(p_rpb_download_state(ctx, xo, vo, force, mo, qo, ato, mfo, mno, mto, &hyd));
  ok = ok && vo[0] > 0.0 && vo[3] < 0.0 && fabs(mass[0] * vo[0] + mass[1] * vo[3]) < 1e-9;
  p_rpb_destroy(ctx);
  printf(ok ? "abi driver: OK\n" : "abi driver: FAILED\n");
  return ok ? 0 : 4;
}
