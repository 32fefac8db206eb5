!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
! BGC_mod - drop-in replacement of the reference module of the same name
! (E3SM-Project/Ocean-BGC, BGC_mod.F90).  Same module name, same public
! entities, same procedure signatures and the same derived types (BGC_parms is
! the reference's own file, compiled unchanged); the bodies forward to the
! B200 library through ISO_C_BINDING:
!
!   BGC_SourceSink     (ref. BGC_mod.F90:340-1998)  -> bgc_source_sink
!   BGC_SurfaceFluxes  (ref. BGC_mod.F90:2706-2957) -> bgc_surface_fluxes
!   BGC_init           (ref. BGC_mod.F90:184-333)   host-side metadata + index wiring
!
! The numerical work - the column_loop, init/compute_particulate_terms and the
! co2calc carbonate solve - runs in CUDA kernels; this file holds none of it.
! Arrays are handed over in place (c_loc of the allocatable components, level
! index fastest as the reference stores them); the library stages them through
! its device arena and returns when the results are back in host memory.
!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
module BGC_mod
  use, intrinsic :: iso_c_binding
  use BGC_parms
  use bgc_b200_capi
  use bgc_b200_runtime
  implicit none
  private

  public :: BGC_tracer_cnt, BGC_init, BGC_SurfaceFluxes, BGC_SourceSink

  integer (BGC_i4), parameter :: BGC_tracer_cnt = 30

  ! nutrient restoring switches: private and never assigned in the reference
  ! (BGC_mod.F90:131-134), i.e. .false.; kept so that a host can flip them here
  logical (BGC_log) :: lrest_po4 = .false., lrest_no3 = .false., lrest_sio3 = .false.

  ! BGC_SurfaceFluxes does not receive the autotroph table, yet the library validates the
  ! tracer-slot map as a whole: remember the table last seen by BGC_init / BGC_SourceSink
  type(autotroph_type), dimension(autotroph_cnt), save :: last_autotrophs
  logical, save :: have_autotrophs = .false.

contains

!-----------------------------------------------------------------------
  subroutine BGC_init(BGC_indices, autotrophs)
    type(autotroph_type), dimension(autotroph_cnt), intent(inout) :: autotrophs
    type(BGC_indices_type), intent(inout) :: BGC_indices
    integer (BGC_i4) :: a, iChl, iC, iFe

    call meta(BGC_indices%po4_ind,  'PO4',  'Dissolved Inorganic Phosphate')
    call meta(BGC_indices%no3_ind,  'NO3',  'Dissolved Inorganic Nitrate')
    call meta(BGC_indices%sio3_ind, 'SiO3', 'Dissolved Inorganic Silicate')
    call meta(BGC_indices%nh4_ind,  'NH4',  'Dissolved Ammonia')
    call meta(BGC_indices%fe_ind,   'Fe',   'Dissolved Inorganic Iron')
    call meta(BGC_indices%o2_ind,   'O2',   'Dissolved Oxygen')
    call meta(BGC_indices%dic_ind,  'DIC',  'Dissolved Inorganic Carbon')
    call meta(BGC_indices%dic_alt_co2_ind, 'DIC_ALT_CO2', 'Dissolved Inorganic Carbon, Alternative CO2')
    call meta(BGC_indices%alk_ind,  'ALK',  'Alkalinity')
    call meta(BGC_indices%doc_ind,  'DOC',  'Dissolved Organic Carbon')
    call meta(BGC_indices%don_ind,  'DON',  'Dissolved Organic Nitrogen')
    call meta(BGC_indices%dofe_ind, 'DOFe', 'Dissolved Organic Iron')
    call meta(BGC_indices%dop_ind,  'DOP',  'Dissolved Organic Phosphorus')
    call meta(BGC_indices%dopr_ind, 'DOPr', 'Refractory DOP')
    call meta(BGC_indices%donr_ind, 'DONr', 'Refractory DON')
    call meta(BGC_indices%zooC_ind, 'zooC', 'Zooplankton Carbon')

    ! functional groups: tracer slots by group, names from the group's own sname/lname
    do a = 1, autotroph_cnt
      if (a == BGC_indices%sp_ind) then
        iChl = BGC_indices%spChl_ind;    iC = BGC_indices%spC_ind;    iFe = BGC_indices%spFe_ind
      else if (a == BGC_indices%diat_ind) then
        iChl = BGC_indices%diatChl_ind;  iC = BGC_indices%diatC_ind;  iFe = BGC_indices%diatFe_ind
      else if (a == BGC_indices%diaz_ind) then
        iChl = BGC_indices%diazChl_ind;  iC = BGC_indices%diazC_ind;  iFe = BGC_indices%diazFe_ind
      else if (a == BGC_indices%phaeo_ind) then
        iChl = BGC_indices%phaeoChl_ind; iC = BGC_indices%phaeoC_ind; iFe = BGC_indices%phaeoFe_ind
      else
        cycle
      end if
      call meta(iChl, trim(autotrophs(a)%sname)//'Chl', trim(autotrophs(a)%lname)//' Chlorophyll')
      call meta(iC,   trim(autotrophs(a)%sname)//'C',   trim(autotrophs(a)%lname)//' Carbon')
      call meta(iFe,  trim(autotrophs(a)%sname)//'Fe',  trim(autotrophs(a)%lname)//' Iron')
      autotrophs(a)%Chl_ind = iChl
      autotrophs(a)%C_ind   = iC
      autotrophs(a)%Fe_ind  = iFe
      autotrophs(a)%Si_ind  = 0
      if (autotrophs(a)%kSiO3 > 0.0_BGC_r8) then
        autotrophs(a)%Si_ind = BGC_indices%diatSi_ind
        call meta(BGC_indices%diatSi_ind, trim(autotrophs(a)%sname)//'Si', trim(autotrophs(a)%lname)//' Silicon')
      end if
      autotrophs(a)%CaCO3_ind = 0
      if (autotrophs(a)%imp_calcifier .or. autotrophs(a)%exp_calcifier) then
        autotrophs(a)%CaCO3_ind = BGC_indices%spCaCO3_ind
        call meta(BGC_indices%spCaCO3_ind, trim(autotrophs(a)%sname)//'CaCO3', trim(autotrophs(a)%lname)//' CaCO3')
      end if
    end do

    BGC_indices%units(:) = 'mmol/m^3'
    BGC_indices%units(BGC_indices%alk_ind) = 'meq/m^3'
    BGC_indices%units(BGC_indices%spChl_ind) = 'mg/m^3'
    BGC_indices%units(BGC_indices%diatChl_ind) = 'mg/m^3'
    BGC_indices%units(BGC_indices%diazChl_ind) = 'mg/m^3'
    BGC_indices%units(BGC_indices%phaeoChl_ind) = 'mg/m^3'

    last_autotrophs = autotrophs
    have_autotrophs = .true.

  contains
    subroutine meta(ind, sname, lname)
      integer (BGC_i4), intent(in) :: ind
      character(len=*), intent(in) :: sname, lname
      BGC_indices%short_name(ind) = sname
      BGC_indices%long_name(ind) = lname
    end subroutine meta
  end subroutine BGC_init

!-----------------------------------------------------------------------
! Flatten the parameter tables (module variables of BGC_parms, the autotroph
! records and the tracer slots) into the C blocks and hand them to the ctx.
! bgc_set_params re-uploads the __constant__ tables only when a value changed.
  subroutine push_params(ctx, autotrophs, BGC_indices)
    type(c_ptr), intent(in) :: ctx
    type(autotroph_type), dimension(autotroph_cnt), intent(in) :: autotrophs
    type(BGC_indices_type), intent(in) :: BGC_indices
    type(BgcParams) :: p
    type(BgcAutotroph) :: ca(4)
    type(BgcIndices) :: ci
    integer :: a

    p%parm_Fe_bioavail = parm_Fe_bioavail;       p%parm_o2_min = parm_o2_min
    p%parm_o2_min_delta = parm_o2_min_delta;     p%parm_kappa_nitrif = parm_kappa_nitrif
    p%parm_nitrif_par_lim = parm_nitrif_par_lim; p%parm_z_mort_0 = parm_z_mort_0
    p%parm_z_mort2_0 = parm_z_mort2_0;           p%parm_labile_ratio = parm_labile_ratio
    p%parm_POMbury = parm_POMbury;               p%parm_BSIbury = parm_BSIbury
    p%parm_fe_scavenge_rate0 = parm_fe_scavenge_rate0
    p%parm_f_prod_sp_CaCO3 = parm_f_prod_sp_CaCO3
    p%parm_POC_diss = parm_POC_diss;             p%parm_SiO2_diss = parm_SiO2_diss
    p%parm_CaCO3_diss = parm_CaCO3_diss
    p%parm_scalelen_z = parm_scalelen_z;         p%parm_scalelen_vals = parm_scalelen_vals
    p%T0_Kelvin_BGC = T0_Kelvin_BGC
    ! whatever value the Fortran compiler gave these literals (single-precision
    ! constants widened, or exact under -fdefault-real-8) is what the GPU uses
    p%epsC = epsC;  p%epsTinv = epsTinv;  p%epsnondim = epsnondim
    p%dust_fescav_scale = dust_fescav_scale;  p%cks = cks;  p%cksi = cksi
    p%lrest_po4 = merge(1_c_int, 0_c_int, lrest_po4)
    p%lrest_no3 = merge(1_c_int, 0_c_int, lrest_no3)
    p%lrest_sio3 = merge(1_c_int, 0_c_int, lrest_sio3)
    p%reserved = 0_c_int

    do a = 1, autotroph_cnt
      ca(a)%Nfixer = merge(1_c_int, 0_c_int, autotrophs(a)%Nfixer)
      ca(a)%imp_calcifier = merge(1_c_int, 0_c_int, autotrophs(a)%imp_calcifier)
      ca(a)%exp_calcifier = merge(1_c_int, 0_c_int, autotrophs(a)%exp_calcifier)
      ca(a)%grazee_ind = autotrophs(a)%grazee_ind;   ca(a)%temp_function = autotrophs(a)%temp_function
      ca(a)%Chl_ind = autotrophs(a)%Chl_ind;  ca(a)%C_ind = autotrophs(a)%C_ind;  ca(a)%Fe_ind = autotrophs(a)%Fe_ind
      ca(a)%Si_ind = autotrophs(a)%Si_ind;    ca(a)%CaCO3_ind = autotrophs(a)%CaCO3_ind
      ca(a)%kFe = autotrophs(a)%kFe;    ca(a)%kPO4 = autotrophs(a)%kPO4;  ca(a)%kDOP = autotrophs(a)%kDOP
      ca(a)%kNO3 = autotrophs(a)%kNO3;  ca(a)%kNH4 = autotrophs(a)%kNH4;  ca(a)%kSiO3 = autotrophs(a)%kSiO3
      ca(a)%Qp = autotrophs(a)%Qp;      ca(a)%gQfe_0 = autotrophs(a)%gQfe_0
      ca(a)%gQfe_min = autotrophs(a)%gQfe_min;       ca(a)%alphaPI = autotrophs(a)%alphaPI
      ca(a)%PCref = autotrophs(a)%PCref;             ca(a)%thetaN_max = autotrophs(a)%thetaN_max
      ca(a)%loss_thres = autotrophs(a)%loss_thres;   ca(a)%loss_thres2 = autotrophs(a)%loss_thres2
      ca(a)%temp_thres = autotrophs(a)%temp_thres;   ca(a)%temp_thresS = autotrophs(a)%temp_thresS
      ca(a)%temp_thresN = autotrophs(a)%temp_thresN; ca(a)%temp_optN = autotrophs(a)%temp_optN
      ca(a)%temp_optS = autotrophs(a)%temp_optS;     ca(a)%mort = autotrophs(a)%mort
      ca(a)%mort2 = autotrophs(a)%mort2;             ca(a)%agg_rate_max = autotrophs(a)%agg_rate_max
      ca(a)%agg_rate_min = autotrophs(a)%agg_rate_min; ca(a)%z_umax_0 = autotrophs(a)%z_umax_0
      ca(a)%z_grz = autotrophs(a)%z_grz;             ca(a)%graze_zoo = autotrophs(a)%graze_zoo
      ca(a)%graze_poc = autotrophs(a)%graze_poc;     ca(a)%graze_doc = autotrophs(a)%graze_doc
      ca(a)%loss_poc = autotrophs(a)%loss_poc;       ca(a)%f_zoo_detr = autotrophs(a)%f_zoo_detr
    end do

    ci%po4_ind = BGC_indices%po4_ind;   ci%no3_ind = BGC_indices%no3_ind;   ci%sio3_ind = BGC_indices%sio3_ind
    ci%nh4_ind = BGC_indices%nh4_ind;   ci%fe_ind = BGC_indices%fe_ind;     ci%o2_ind = BGC_indices%o2_ind
    ci%dic_ind = BGC_indices%dic_ind;   ci%dic_alt_co2_ind = BGC_indices%dic_alt_co2_ind
    ci%alk_ind = BGC_indices%alk_ind;   ci%doc_ind = BGC_indices%doc_ind;   ci%don_ind = BGC_indices%don_ind
    ci%dofe_ind = BGC_indices%dofe_ind; ci%dop_ind = BGC_indices%dop_ind;   ci%dopr_ind = BGC_indices%dopr_ind
    ci%donr_ind = BGC_indices%donr_ind; ci%zooC_ind = BGC_indices%zooC_ind
    ci%spC_ind = BGC_indices%spC_ind;   ci%spChl_ind = BGC_indices%spChl_ind
    ci%spFe_ind = BGC_indices%spFe_ind; ci%spCaCO3_ind = BGC_indices%spCaCO3_ind
    ci%diatC_ind = BGC_indices%diatC_ind;   ci%diatChl_ind = BGC_indices%diatChl_ind
    ci%diatFe_ind = BGC_indices%diatFe_ind; ci%diatSi_ind = BGC_indices%diatSi_ind
    ci%phaeoC_ind = BGC_indices%phaeoC_ind; ci%phaeoChl_ind = BGC_indices%phaeoChl_ind
    ci%phaeoFe_ind = BGC_indices%phaeoFe_ind
    ci%diazC_ind = BGC_indices%diazC_ind;   ci%diazChl_ind = BGC_indices%diazChl_ind
    ci%diazFe_ind = BGC_indices%diazFe_ind
    ci%sp_ind = BGC_indices%sp_ind;     ci%diat_ind = BGC_indices%diat_ind
    ci%diaz_ind = BGC_indices%diaz_ind; ci%phaeo_ind = BGC_indices%phaeo_ind

    call bgc_b200_check(bgc_set_params(ctx, p, ca, ci), 'bgc_set_params')
  end subroutine push_params

  subroutine fill_input(BGC_input, cin)
    type(BGC_input_type), intent(in), target :: BGC_input
    type(BgcInput), intent(out) :: cin
    cin%BGC_tracers = loc3(BGC_input%BGC_tracers)
    cin%PotentialTemperature = loc2(BGC_input%PotentialTemperature)
    cin%Salinity = loc2(BGC_input%Salinity)
    cin%cell_center_depth = loc2(BGC_input%cell_center_depth)
    cin%cell_thickness = loc2(BGC_input%cell_thickness)
    cin%cell_bottom_depth = loc2(BGC_input%cell_bottom_depth)
    cin%cell_latitude = loc1(BGC_input%cell_latitude)
    cin%number_of_active_levels = loci1(BGC_input%number_of_active_levels)
  end subroutine fill_input

  subroutine fill_forcing(BGC_forcing, cfo)
    type(BGC_forcing_type), intent(in), target :: BGC_forcing
    type(BgcForcing), intent(out) :: cfo
    cfo%FESEDFLUX = loc2(BGC_forcing%FESEDFLUX)
    cfo%NUTR_RESTORE_RTAU = loc2(BGC_forcing%NUTR_RESTORE_RTAU)
    cfo%NO3_CLIM = loc2(BGC_forcing%NO3_CLIM)
    cfo%PO4_CLIM = loc2(BGC_forcing%PO4_CLIM)
    cfo%SiO3_CLIM = loc2(BGC_forcing%SiO3_CLIM)
    cfo%dust_FLUX_IN = loc1(BGC_forcing%dust_FLUX_IN)
    cfo%ShortWaveFlux_surface = loc1(BGC_forcing%ShortWaveFlux_surface)
    cfo%surfacePressure = loc1(BGC_forcing%surfacePressure)
    cfo%iceFraction = loc1(BGC_forcing%iceFraction)
    cfo%windSpeedSquared10m = loc1(BGC_forcing%windSpeedSquared10m)
    cfo%atmCO2 = loc1(BGC_forcing%atmCO2)
    cfo%atmCO2_ALT_CO2 = loc1(BGC_forcing%atmCO2_ALT_CO2)
    cfo%surface_pH = loc1(BGC_forcing%surface_pH)
    cfo%surface_pH_alt_co2 = loc1(BGC_forcing%surface_pH_alt_co2)
    cfo%surfaceDepth = loc1(BGC_forcing%surfaceDepth)
    cfo%SST = loc1(BGC_forcing%SST)
    cfo%SSS = loc1(BGC_forcing%SSS)
    cfo%depositionFlux = loc2(BGC_forcing%depositionFlux)
    cfo%riverFlux = loc2(BGC_forcing%riverFlux)
    cfo%gasFlux = loc2(BGC_forcing%gasFlux)
    cfo%seaIceFlux = loc2(BGC_forcing%seaIceFlux)
    cfo%netFlux = loc2(BGC_forcing%netFlux)
    cfo%lcalc_O2_gas_flux = merge(1_c_int, 0_c_int, BGC_forcing%lcalc_O2_gas_flux)
    cfo%lcalc_CO2_gas_flux = merge(1_c_int, 0_c_int, BGC_forcing%lcalc_CO2_gas_flux)
  end subroutine fill_forcing

!-----------------------------------------------------------------------
  subroutine BGC_SourceSink(autotrophs, BGC_indices, BGC_input, BGC_forcing,   &
                            BGC_output, BGC_diagnostic_fields, numLevelsMax,   &
                            numColumnsMax, numColumns, alt_co2_use_eco)
    type(autotroph_type), dimension(autotroph_cnt), intent(in) :: autotrophs
    type(BGC_indices_type),     intent(in )          :: BGC_indices
    type(BGC_input_type),       intent(in ), target  :: BGC_input
    type(BGC_forcing_type),     intent(in ), target  :: BGC_forcing
    integer (BGC_i4), intent(in) :: numLevelsMax, numColumnsMax, numColumns
    logical (BGC_log), intent(in) :: alt_co2_use_eco
    type(BGC_output_type),      intent(inout), target :: BGC_output
    type(BGC_diagnostics_type), intent(inout), target :: BGC_diagnostic_fields

    type(c_ptr) :: ctx
    logical :: fresh
    type(BgcInput) :: cin
    type(BgcForcing) :: cfo
    type(BgcOutput) :: cout
    type(BgcDiagnostics) :: cdg

    ctx = bgc_b200_ctx(numLevelsMax, numColumnsMax, fresh)
    call push_params(ctx, autotrophs, BGC_indices)
    last_autotrophs = autotrophs
    have_autotrophs = .true.
    call fill_input(BGC_input, cin)
    call fill_forcing(BGC_forcing, cfo)
    cout%BGC_tendencies = loc3(BGC_output%BGC_tendencies)
    cout%PH_PREV_3D = loc2(BGC_output%PH_PREV_3D)
    cout%PH_PREV_ALT_CO2_3D = loc2(BGC_output%PH_PREV_ALT_CO2_3D)
    include 'bgc_diag_ptrs.inc'

    call bgc_b200_check(bgc_source_sink(ctx, cin, cfo, cout, cdg, int(numLevelsMax, c_int),         &
                                        int(numColumnsMax, c_int), int(numColumns, c_int),          &
                                        merge(1_c_int, 0_c_int, alt_co2_use_eco),                   &
                                        BGC_MEM_HOST_FORTRAN), 'bgc_source_sink')
  end subroutine BGC_SourceSink

!-----------------------------------------------------------------------
  subroutine BGC_SurfaceFluxes(BGC_indices, BGC_input, BGC_forcing,   &
                               BGC_flux_diagnostic_fields,           &
                               numColumnsMax, numColumns)
    type(BGC_indices_type), intent(in )         :: BGC_indices
    type(BGC_input_type),   intent(in ), target :: BGC_input
    type(BGC_forcing_type), intent(inout), target :: BGC_forcing
    integer (BGC_i4), intent(in) :: numColumnsMax, numColumns
    type(BGC_flux_diagnostics_type), intent(inout), target :: BGC_flux_diagnostic_fields

    type(c_ptr) :: ctx
    logical :: fresh
    type(BgcInput) :: cin
    type(BgcForcing) :: cfo
    type(BgcFluxDiagnostics) :: cfd
    integer(c_int) :: nLevelsMax

    ! the tracer array's leading extent is the only place the level count appears here
    nLevelsMax = int(size(BGC_input%BGC_tracers, 1), c_int)
    ctx = bgc_b200_ctx(int(nLevelsMax), numColumnsMax, fresh)
    if (.not. have_autotrophs) error stop 'BGC_SurfaceFluxes: call BGC_init (or BGC_SourceSink) first'
    call push_params(ctx, last_autotrophs, BGC_indices)
    call fill_input(BGC_input, cin)
    call fill_forcing(BGC_forcing, cfo)
    include 'bgc_flux_diag_ptrs.inc'
    call bgc_b200_check(bgc_surface_fluxes(ctx, cin, cfo, cfd, nLevelsMax, int(numColumnsMax, c_int),   &
                                           int(numColumns, c_int), BGC_MEM_HOST_FORTRAN), 'bgc_surface_fluxes')
  end subroutine BGC_SurfaceFluxes

end module BGC_mod
