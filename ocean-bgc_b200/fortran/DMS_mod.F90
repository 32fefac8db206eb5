!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
! DMS_mod - drop-in replacement of the reference module of the same name
! (DMS_mod.F90): same public entities and signatures, DMS_parms unchanged;
! bodies forward to the B200 library.
!
!   DMS_SourceSink     (ref. DMS_mod.F90:156-770) -> dms_source_sink
!   DMS_SurfaceFluxes  (ref. DMS_mod.F90:778-908) -> dms_surface_fluxes
!   DMS_init           (ref. DMS_mod.F90:73-148)  host-side metadata
!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
module DMS_mod
  use, intrinsic :: iso_c_binding
  use DMS_parms
  use bgc_b200_capi
  use bgc_b200_runtime
  implicit none
  private

  public :: DMS_tracer_cnt, DMS_init, DMS_SurfaceFluxes, DMS_SourceSink

  integer (DMS_i4), parameter :: DMS_tracer_cnt = 14

contains

  subroutine DMS_init(DMS_indices)
    type(DMS_indices_type), intent(inout) :: DMS_indices
    call meta(DMS_indices%dms_ind,      'DMS',      'DiMethyl Sulfide')
    call meta(DMS_indices%dmsp_ind,     'DMSP',     'Dimethylsulfoniopropionate')
    call meta(DMS_indices%no3_ind,      'NO3',      'Dissolved Inorganic Nitrate')
    call meta(DMS_indices%doc_ind,      'DOC',      'Dissolved Organic Carbon')
    call meta(DMS_indices%zooC_ind,     'zooC',     'Zooplankton Carbon')
    call meta(DMS_indices%spChl_ind,    'spChl',    ' Small Phytoplankton Chlorophyll')
    call meta(DMS_indices%diatChl_ind,  'diatChl',  ' Diatom Chlorophyll')
    call meta(DMS_indices%diazChl_ind,  'diazChl',  ' Diazotroph Chlorophyll')
    call meta(DMS_indices%phaeoChl_ind, 'phaeoChl', 'Phaeocystis Chlorophyll')
    call meta(DMS_indices%spC_ind,      'spC',      ' Small Phytoplankton Carbon')
    call meta(DMS_indices%diatC_ind,    'diatC',    ' Diatom Carbon')
    call meta(DMS_indices%diazC_ind,    'diazC',    ' Diazotroph Carbon')
    call meta(DMS_indices%phaeoC_ind,   'phaeoC',   'Phaeocystis Carbon')
    call meta(DMS_indices%spCaCO3_ind,  'spCaCO3',  ' Small Phytoplankton Calcium Carbonate')
    DMS_indices%units(:) = 'mmol/m^3'
  contains
    subroutine meta(ind, sname, lname)
      integer (DMS_i4), intent(in) :: ind
      character(len=*), intent(in) :: sname, lname
      DMS_indices%short_name(ind) = sname
      DMS_indices%long_name(ind) = lname
    end subroutine meta
  end subroutine DMS_init

  subroutine push_params(ctx, DMS_indices)
    type(c_ptr), intent(in) :: ctx
    type(DMS_indices_type), intent(in) :: DMS_indices
    type(DmsParams) :: p
    type(DmsIndices) :: ci
    p%k_S_p_base = k_S_p_base;  p%zooC_avg = zooC_avg;  p%mort = mort;  p%k_conv = k_conv
    p%k_S_z = k_S_z;  p%B_preexp = B_preexp;  p%B_exp = B_exp;  p%k_S_B = k_S_B;  p%k_bkgnd = k_bkgnd
    p%j_dms_perI = j_dms_perI;  p%inject_scale = inject_scale
    p%T_cryo_hi = T_cryo_hi;  p%T_cryo_lo = T_cryo_lo;  p%T_lo = T_lo;  p%T_hi = T_hi
    p%Min_cyano_frac = Min_cyano_frac;  p%Max_cyano_frac = Max_cyano_frac
    p%Min_yld = Min_yld;  p%Max_yld = Max_yld;  p%G_phaeo_S = G_phaeo_S;  p%Sp_ref = Sp_ref
    p%Stress_mult = Stress_mult;  p%R = R
    p%Rs2n_diat = Rs2n_diat;  p%Rs2n_phaeo = Rs2n_phaeo;  p%Rs2n_cocco = Rs2n_cocco
    p%Rs2n_cyano = Rs2n_cyano;  p%Rs2n_eukar = Rs2n_eukar;  p%Rs2n_diaz = Rs2n_diaz
    p%f_qsw_par_DMS = f_qsw_par_DMS
    ci%dms_ind = DMS_indices%dms_ind;      ci%dmsp_ind = DMS_indices%dmsp_ind
    ci%no3_ind = DMS_indices%no3_ind;      ci%doc_ind = DMS_indices%doc_ind
    ci%zooC_ind = DMS_indices%zooC_ind;    ci%spC_ind = DMS_indices%spC_ind
    ci%spCaCO3_ind = DMS_indices%spCaCO3_ind
    ci%diatC_ind = DMS_indices%diatC_ind;  ci%diazC_ind = DMS_indices%diazC_ind
    ci%phaeoC_ind = DMS_indices%phaeoC_ind; ci%spChl_ind = DMS_indices%spChl_ind
    ci%diatChl_ind = DMS_indices%diatChl_ind; ci%diazChl_ind = DMS_indices%diazChl_ind
    ci%phaeoChl_ind = DMS_indices%phaeoChl_ind
    call bgc_b200_check(dms_set_params(ctx, p, ci), 'dms_set_params')
  end subroutine push_params

  subroutine fill_input(DMS_input, cin)
    type(DMS_input_type), intent(in), target :: DMS_input
    type(DmsInput), intent(out) :: cin
    cin%DMS_tracers = loc3(DMS_input%DMS_tracers)
    cin%cell_thickness = loc2(DMS_input%cell_thickness)
    cin%number_of_active_levels = loci1(DMS_input%number_of_active_levels)
  end subroutine fill_input

  subroutine fill_forcing(DMS_forcing, cfo)
    type(DMS_forcing_type), intent(in), target :: DMS_forcing
    type(DmsForcing), intent(out) :: cfo
    cfo%ShortWaveFlux_surface = loc1(DMS_forcing%ShortWaveFlux_surface)
    cfo%surfacePressure = loc1(DMS_forcing%surfacePressure)
    cfo%iceFraction = loc1(DMS_forcing%iceFraction)
    cfo%windSpeedSquared10m = loc1(DMS_forcing%windSpeedSquared10m)
    cfo%SST = loc1(DMS_forcing%SST)
    cfo%SSS = loc1(DMS_forcing%SSS)
    cfo%netFlux = loc2(DMS_forcing%netFlux)
    cfo%lcalc_DMS_gas_flux = merge(1_c_int, 0_c_int, DMS_forcing%lcalc_DMS_gas_flux)
  end subroutine fill_forcing

  subroutine DMS_SourceSink(DMS_indices, DMS_input, DMS_forcing, DMS_output, DMS_diagnostic_fields, &
                            numLevelsMax, numColumnsMax, numColumns)
    type(DMS_indices_type),     intent(in )           :: DMS_indices
    type(DMS_input_type),       intent(in ), target   :: DMS_input
    type(DMS_forcing_type),     intent(in ), target   :: DMS_forcing
    integer (DMS_i4), intent(in) :: numLevelsMax, numColumnsMax, numColumns
    type(DMS_output_type),      intent(inout), target :: DMS_output
    type(DMS_diagnostics_type), intent(inout), target :: DMS_diagnostic_fields
    type(c_ptr) :: ctx
    logical :: fresh
    type(DmsInput) :: cin
    type(DmsForcing) :: cfo
    type(DmsOutput) :: cout
    type(DmsDiagnostics) :: cdg

    ctx = bgc_b200_ctx(numLevelsMax, numColumnsMax, fresh)
    call push_params(ctx, DMS_indices)
    call fill_input(DMS_input, cin)
    call fill_forcing(DMS_forcing, cfo)
    cout%DMS_tendencies = loc3(DMS_output%DMS_tendencies)
    include 'dms_diag_ptrs.inc'
    call bgc_b200_check(dms_source_sink(ctx, cin, cfo, cout, cdg, int(numLevelsMax, c_int),      &
                                        int(numColumnsMax, c_int), int(numColumns, c_int),       &
                                        BGC_MEM_HOST_FORTRAN), 'dms_source_sink')
  end subroutine DMS_SourceSink

  subroutine DMS_SurfaceFluxes(DMS_indices, DMS_input, DMS_forcing,   &
                               DMS_flux_diagnostic_fields, numColumnsMax, numColumns)
    type(DMS_indices_type), intent(in )           :: DMS_indices
    type(DMS_input_type),   intent(in ), target   :: DMS_input
    type(DMS_forcing_type), intent(inout), target :: DMS_forcing
    integer (DMS_i4), intent(in) :: numColumnsMax, numColumns
    type(DMS_flux_diagnostics_type), intent(inout), target :: DMS_flux_diagnostic_fields
    type(c_ptr) :: ctx
    logical :: fresh
    type(DmsInput) :: cin
    type(DmsForcing) :: cfo
    type(DmsFluxDiagnostics) :: cfd
    integer(c_int) :: nLevelsMax

    nLevelsMax = int(size(DMS_input%DMS_tracers, 1), c_int)
    ctx = bgc_b200_ctx(int(nLevelsMax), numColumnsMax, fresh)
    call push_params(ctx, DMS_indices)
    call fill_input(DMS_input, cin)
    call fill_forcing(DMS_forcing, cfo)
    include 'dms_flux_diag_ptrs.inc'
    call bgc_b200_check(dms_surface_fluxes(ctx, cin, cfo, cfd, nLevelsMax, int(numColumnsMax, c_int),   &
                                           int(numColumns, c_int), BGC_MEM_HOST_FORTRAN), 'dms_surface_fluxes')
  end subroutine DMS_SurfaceFluxes

end module DMS_mod
