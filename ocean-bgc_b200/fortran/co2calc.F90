!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
! co2calc - drop-in replacement of the reference module of the same name
! (co2calc.F90:24: PUBLIC :: co2calc_1point, comp_CO3terms, comp_co3_sat_vals).
! Same public entities and argument lists; every call runs on the GPU through
! the C ABI (there is no CPU code path in this library):
!
!   co2calc_1point     (ref. co2calc.F90:75-210)    -> bgc_co2calc_points     (n = 1)
!   comp_CO3terms      (ref. co2calc.F90:214-316)   -> bgc_comp_co3terms      (n = 1)
!   comp_co3_sat_vals  (ref. co2calc.F90:1096-1238) -> bgc_comp_co3_sat_vals  (n = 1)
!
! A scalar call costs a kernel launch and a round trip to the device; hosts that
! have many points should use the *_points forms below (extensions: same
! arguments as arrays of length n, one launch for all of them).  BGC_SourceSink /
! BGC_SurfaceFluxes of this library do not go through this module: their carbonate
! solves are fused into their own kernels.
!
! lcomp_co3_coeffs = .false. asks the reference to reuse the equilibrium constants
! its previous call left in module SAVE variables (co2calc.F90:65-67).  This
! library keeps no such state: the constants are always computed from the
! arguments of the call, which gives the same answer whenever the reference's
! reuse is legitimate (same k, depth, temp, salt as the previous call).
! locmip_k1_k2_bug_fix is accepted and ignored exactly as in the reference (the
! dummy argument is never read, co2calc.F90:75-210).
!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
module co2calc
  use, intrinsic :: iso_c_binding
  use BGC_parms
  use bgc_b200_capi
  use bgc_b200_runtime
  implicit none
  private

  public :: co2calc_1point, comp_CO3terms, comp_co3_sat_vals
  public :: co2calc_points, comp_CO3terms_points, comp_co3_sat_vals_points

contains

  subroutine co2calc_1point(depth, locmip_k1_k2_bug_fix, lcomp_co3_coeffs, &
       temp, salt, dic_in, ta_in, pt_in, sit_in, phlo, phhi, ph, xco2_in, atmpres, &
       co2star, dco2star, pCO2surf, dpco2)
    logical (BGC_log), intent(in) :: locmip_k1_k2_bug_fix
    logical (BGC_log), intent(in) :: lcomp_co3_coeffs
    real (BGC_r8), intent(in) :: depth, temp, salt, dic_in, ta_in, pt_in, sit_in, xco2_in, atmpres
    real (BGC_r8), intent(inout) :: phlo, phhi
    real (BGC_r8), intent(out) :: ph, co2star, dco2star, pCO2surf, dpco2
    real (c_double) :: a_depth(1), a_temp(1), a_salt(1), a_dic(1), a_ta(1), a_pt(1), a_sit(1), a_lo(1), a_hi(1), &
                       a_xco2(1), a_atm(1), o_ph(1), o_co2star(1), o_dco2star(1), o_pco2(1), o_dpco2(1)
    a_depth(1) = depth; a_temp(1) = temp; a_salt(1) = salt; a_dic(1) = dic_in; a_ta(1) = ta_in
    a_pt(1) = pt_in; a_sit(1) = sit_in; a_lo(1) = phlo; a_hi(1) = phhi; a_xco2(1) = xco2_in; a_atm(1) = atmpres
    call co2calc_points(1, a_depth, a_temp, a_salt, a_dic, a_ta, a_pt, a_sit, a_lo, a_hi, a_xco2, a_atm, &
                        o_ph, o_co2star, o_dco2star, o_pco2, o_dpco2)
    ph = o_ph(1); co2star = o_co2star(1); dco2star = o_dco2star(1); pCO2surf = o_pco2(1); dpco2 = o_dpco2(1)
  end subroutine co2calc_1point

  subroutine comp_CO3terms(k, depth, lcomp_co3_coeffs, temp, salt, &
       dic_in, ta_in, pt_in, sit_in, phlo, phhi, ph, H2CO3, HCO3, CO3)
    integer (BGC_i4), intent(in) :: k
    logical (BGC_log), intent(in) :: lcomp_co3_coeffs
    real (BGC_r8), intent(in) :: depth, temp, salt, dic_in, ta_in, pt_in, sit_in
    real (BGC_r8), intent(inout) :: phlo, phhi
    real (BGC_r8), intent(out) :: ph, H2CO3, HCO3, CO3
    integer (c_int) :: a_k(1)
    real (c_double) :: a_depth(1), a_temp(1), a_salt(1), a_dic(1), a_ta(1), a_pt(1), a_sit(1), a_lo(1), a_hi(1), &
                       o_ph(1), o_h2co3(1), o_hco3(1), o_co3(1)
    a_k(1) = int(k, c_int)
    a_depth(1) = depth; a_temp(1) = temp; a_salt(1) = salt; a_dic(1) = dic_in; a_ta(1) = ta_in
    a_pt(1) = pt_in; a_sit(1) = sit_in; a_lo(1) = phlo; a_hi(1) = phhi
    call comp_CO3terms_points(1, a_k, a_depth, a_temp, a_salt, a_dic, a_ta, a_pt, a_sit, a_lo, a_hi, &
                              o_ph, o_h2co3, o_hco3, o_co3)
    ph = o_ph(1); H2CO3 = o_h2co3(1); HCO3 = o_hco3(1); CO3 = o_co3(1)
  end subroutine comp_CO3terms

  subroutine comp_co3_sat_vals(k, depth, temp, salt, co3_sat_calc, co3_sat_arag)
    integer (BGC_i4), intent(in) :: k
    real (BGC_r8), intent(in) :: depth, temp, salt
    real (BGC_r8), intent(out) :: co3_sat_calc, co3_sat_arag
    integer (c_int) :: a_k(1)
    real (c_double) :: a_depth(1), a_temp(1), a_salt(1), o_calc(1), o_arag(1)
    a_k(1) = int(k, c_int)
    a_depth(1) = depth; a_temp(1) = temp; a_salt(1) = salt
    call comp_co3_sat_vals_points(1, a_k, a_depth, a_temp, a_salt, o_calc, o_arag)
    co3_sat_calc = o_calc(1); co3_sat_arag = o_arag(1)
  end subroutine comp_co3_sat_vals

  ! ---- batched forms (extensions): arrays of length n, one kernel launch
  subroutine co2calc_points(n, depth, temp, salt, dic_in, ta_in, pt_in, sit_in, phlo, phhi, xco2_in, atmpres, &
                            ph, co2star, dco2star, pCO2surf, dpco2)
    integer, intent(in) :: n
    real (c_double), intent(in) :: depth(*), temp(*), salt(*), dic_in(*), ta_in(*), pt_in(*), sit_in(*), &
                                   phlo(*), phhi(*), xco2_in(*), atmpres(*)
    real (c_double), intent(out) :: ph(*), co2star(*), dco2star(*), pCO2surf(*), dpco2(*)
    type(c_ptr) :: ctx
    logical :: fresh
    ctx = bgc_b200_ctx(1, 1, fresh)
    call bgc_b200_check(bgc_co2calc_points(ctx, int(n, c_int), depth, temp, salt, dic_in, ta_in, pt_in, sit_in, &
                                           phlo, phhi, xco2_in, atmpres, ph, co2star, dco2star, pCO2surf, dpco2, &
                                           BGC_MEM_HOST_FORTRAN), 'bgc_co2calc_points')
  end subroutine co2calc_points

  subroutine comp_CO3terms_points(n, k, depth, temp, salt, dic_in, ta_in, pt_in, sit_in, phlo, phhi, &
                                  ph, H2CO3, HCO3, CO3)
    integer, intent(in) :: n
    integer (c_int), intent(in) :: k(*)
    real (c_double), intent(in) :: depth(*), temp(*), salt(*), dic_in(*), ta_in(*), pt_in(*), sit_in(*), phlo(*), phhi(*)
    real (c_double), intent(out) :: ph(*), H2CO3(*), HCO3(*), CO3(*)
    type(c_ptr) :: ctx
    logical :: fresh
    ctx = bgc_b200_ctx(1, 1, fresh)
    call bgc_b200_check(bgc_comp_co3terms(ctx, int(n, c_int), k, int(1, c_int), depth, temp, salt, dic_in, ta_in, pt_in, &
                                          sit_in, phlo, phhi, ph, H2CO3, HCO3, CO3, BGC_MEM_HOST_FORTRAN), &
                        'bgc_comp_co3terms')
  end subroutine comp_CO3terms_points

  subroutine comp_co3_sat_vals_points(n, k, depth, temp, salt, co3_sat_calc, co3_sat_arag)
    integer, intent(in) :: n
    integer (c_int), intent(in) :: k(*)
    real (c_double), intent(in) :: depth(*), temp(*), salt(*)
    real (c_double), intent(out) :: co3_sat_calc(*), co3_sat_arag(*)
    type(c_ptr) :: ctx
    logical :: fresh
    ctx = bgc_b200_ctx(1, 1, fresh)
    call bgc_b200_check(bgc_comp_co3_sat_vals(ctx, int(n, c_int), k, int(1, c_int), depth, temp, salt, co3_sat_calc, &
                                              co3_sat_arag, BGC_MEM_HOST_FORTRAN), 'bgc_comp_co3_sat_vals')
  end subroutine comp_co3_sat_vals_points

end module co2calc
