// Moist (EquilibriumMicrophysics0M) switch of the CTA emulators: emu_set_moist(mp, Hw) with
// mp = [R_v, cp_v, cp_l, cp_i, LH_v0, LH_s0, T_triple, press_triple, T_freeze, T_icenuc, pow_icenuc] turns the following calls of this
// library into moist ones (Par::moist = 1, Par::M filled as capi.cu:make_par does); mp = NULL switches back to dry.  Hw: the
// ρ(h_eff + Φ) buffer [nh][16][nv] that k5_exp_a<…, MOIST> writes and part 1 of k7_exp_c reads.  Test infrastructure only.
#pragma once
static double g_moist_par[11];
static int g_moist_on = 0;
static void* g_moist_Hw = nullptr;
extern "C" __attribute__((visibility("default"))) int emu_set_moist(const double* mp, void* Hw) {
  g_moist_on = mp != nullptr;
  if (mp) memcpy(g_moist_par, mp, sizeof(g_moist_par));
  g_moist_Hw = Hw;
  return 0;
}
template <class P_>
static void emu_apply_moist(P_& P, double T_0) {
  P.moist = g_moist_on;
  if (!g_moist_on) return;
  const double* m = g_moist_par;
  P.M.R_v = m[0]; P.M.cv_v = m[1] - m[0]; P.M.cp_v = m[1]; P.M.cp_l = m[2]; P.M.cp_i = m[3]; P.M.LH_v0 = m[4]; P.M.LH_s0 = m[5];
  P.M.e_v0 = m[4] - m[0] * T_0; P.M.e_i0 = m[5] - m[4]; P.M.T_tr = m[6]; P.M.ln_ptr = std::log(m[7]); P.M.T_frz = m[8]; P.M.T_icn = m[9];
  P.M.pow_icn = m[10];
  const double R_d = P.R_d;
  P.M.A_liq = (m[1] - m[2]) / m[0]; P.M.A_ice = (m[1] - m[3]) / m[0];
  P.M.B_liq = (m[4] - (m[1] - m[2]) * T_0) / m[0]; P.M.B_ice = (m[5] - (m[1] - m[3]) * T_0) / m[0];
  P.M.iT_tr = 1.0 / m[6]; P.M.epsv = m[0] / R_d;
  P.M.q_neg = 0.25 * 2.220446049250313e-16 * P.cv_d / m[5];
}
