!  ref_driver - runs the UNMODIFIED reference (E3SM-Project/Ocean-BGC) on one block of
!  columns read from a flat binary file and dumps every output, so that the CPU oracle of
!  this repository can be pinned against the reference wherever a Fortran compiler exists
!  (tests/fortran/README.md).  Usage: ref_driver <input.bin> <output.bin>
!
!  Input file (stream, little endian, written by tests/test_fortran_reference.py):
!    int32  nL, nC, nCols, alt_co2_use_eco
!    real64 T0_Kelvin_BGC
!    real64 BGC_tracers(nL,nC,30), T, S, zmid, dz, zbot (nL,nC each), lat(nC); int32 kmax(nC)
!    real64 FESEDFLUX(nL,nC), dust(nC), sw(nC), surfacePressure, iceFraction, windSpeedSquared10m,
!           atmCO2, atmCO2_ALT_CO2, surface_pH, surface_pH_alt_co2, surfaceDepth, SST, SSS (nC each),
!           depositionFlux, riverFlux, gasFlux, seaIceFlux, netFlux (nC,30 each)
!    real64 PH_PREV_3D, PH_PREV_ALT_CO2_3D (nL,nC)
!    real64 DMS_tracers(nL,nC,14), MACROS_tracers(nL,nC,8)
!  Output file: see the write statements at the end (same order the test reads them).
program ref_driver
  use BGC_parms
  use BGC_mod
  use DMS_parms
  use DMS_mod
  use MACROS_parms
  use MACROS_mod
  implicit none
  integer, parameter :: iu = 11, ou = 12
  real (BGC_r8), parameter :: fill_value = 7.25_BGC_r8
  character(len=1024) :: fin, fout
  integer (BGC_i4) :: nL, nC, nCols, ialt
  logical (BGC_log) :: alt
  type(autotroph_type), dimension(autotroph_cnt) :: autotrophs
  type(BGC_indices_type) :: bind
  type(BGC_input_type) :: bin
  type(BGC_forcing_type) :: bfo
  type(BGC_output_type) :: bout
  type(BGC_diagnostics_type) :: bdiag
  type(BGC_flux_diagnostics_type) :: bfdiag
  type(DMS_indices_type) :: dind
  type(DMS_input_type) :: din
  type(DMS_forcing_type) :: dfo
  type(DMS_output_type) :: dout
  type(DMS_diagnostics_type) :: ddiag
  type(DMS_flux_diagnostics_type) :: dfdiag
  type(MACROS_indices_type) :: mind
  type(MACROS_input_type) :: min_
  type(MACROS_output_type) :: mout
  type(MACROS_diagnostics_type) :: mdiag
  real (BGC_r8), allocatable :: ph_cold(:,:)

  call get_command_argument(1, fin)
  call get_command_argument(2, fout)
  open(iu, file=trim(fin), access='stream', form='unformatted', status='old')
  read(iu) nL, nC, nCols, ialt
  alt = ialt /= 0
  read(iu) T0_Kelvin_BGC                      ! never assigned by the reference (BGC_parms.F90:45)

  ! ---- tracer slots in declaration order (what the host chooses), default tables
  bind%po4_ind = 1;  bind%no3_ind = 2;  bind%sio3_ind = 3;  bind%nh4_ind = 4;  bind%fe_ind = 5
  bind%o2_ind = 6;   bind%dic_ind = 7;  bind%dic_alt_co2_ind = 8;  bind%alk_ind = 9;  bind%doc_ind = 10
  bind%don_ind = 11; bind%dofe_ind = 12; bind%dop_ind = 13; bind%dopr_ind = 14; bind%donr_ind = 15
  bind%zooC_ind = 16; bind%spC_ind = 17; bind%spChl_ind = 18; bind%spFe_ind = 19; bind%spCaCO3_ind = 20
  bind%diatC_ind = 21; bind%diatChl_ind = 22; bind%diatFe_ind = 23; bind%diatSi_ind = 24
  bind%phaeoC_ind = 25; bind%phaeoChl_ind = 26; bind%phaeoFe_ind = 27
  bind%diazC_ind = 28; bind%diazChl_ind = 29; bind%diazFe_ind = 30
  allocate(bind%short_name(BGC_tracer_cnt), bind%long_name(BGC_tracer_cnt), bind%units(BGC_tracer_cnt))
  call BGC_parms_init(bind, autotrophs)
  call BGC_init(bind, autotrophs)

  ! ---- BGC input / forcing / output
  allocate(bin%BGC_tracers(nL,nC,BGC_tracer_cnt), bin%PotentialTemperature(nL,nC), bin%Salinity(nL,nC), &
           bin%cell_center_depth(nL,nC), bin%cell_thickness(nL,nC), bin%cell_bottom_depth(nL,nC), &
           bin%cell_latitude(nC), bin%number_of_active_levels(nC))
  read(iu) bin%BGC_tracers, bin%PotentialTemperature, bin%Salinity, bin%cell_center_depth, &
           bin%cell_thickness, bin%cell_bottom_depth, bin%cell_latitude, bin%number_of_active_levels
  allocate(bfo%FESEDFLUX(nL,nC), bfo%NUTR_RESTORE_RTAU(nL,nC), bfo%NO3_CLIM(nL,nC), bfo%PO4_CLIM(nL,nC), &
           bfo%SiO3_CLIM(nL,nC), bfo%dust_FLUX_IN(nC), bfo%ShortWaveFlux_surface(nC), bfo%surfacePressure(nC), &
           bfo%iceFraction(nC), bfo%windSpeedSquared10m(nC), bfo%atmCO2(nC), bfo%atmCO2_ALT_CO2(nC), &
           bfo%surface_pH(nC), bfo%surface_pH_alt_co2(nC), bfo%surfaceDepth(nC), bfo%SST(nC), bfo%SSS(nC), &
           bfo%depositionFlux(nC,BGC_tracer_cnt), bfo%riverFlux(nC,BGC_tracer_cnt), bfo%gasFlux(nC,BGC_tracer_cnt), &
           bfo%seaIceFlux(nC,BGC_tracer_cnt), bfo%netFlux(nC,BGC_tracer_cnt))
  bfo%NUTR_RESTORE_RTAU = 0.0_BGC_r8; bfo%NO3_CLIM = 0.0_BGC_r8; bfo%PO4_CLIM = 0.0_BGC_r8; bfo%SiO3_CLIM = 0.0_BGC_r8
  read(iu) bfo%FESEDFLUX, bfo%dust_FLUX_IN, bfo%ShortWaveFlux_surface, bfo%surfacePressure, bfo%iceFraction, &
           bfo%windSpeedSquared10m, bfo%atmCO2, bfo%atmCO2_ALT_CO2, bfo%surface_pH, bfo%surface_pH_alt_co2, &
           bfo%surfaceDepth, bfo%SST, bfo%SSS, bfo%depositionFlux, bfo%riverFlux, bfo%gasFlux, bfo%seaIceFlux, &
           bfo%netFlux
  bfo%lcalc_O2_gas_flux = .true.
  bfo%lcalc_CO2_gas_flux = .true.
  allocate(bout%BGC_tendencies(nL,nC,BGC_tracer_cnt), bout%PH_PREV_3D(nL,nC), bout%PH_PREV_ALT_CO2_3D(nL,nC), ph_cold(nL,nC))
  bout%BGC_tendencies = fill_value
  read(iu) bout%PH_PREV_3D, bout%PH_PREV_ALT_CO2_3D

  ! ---- DMS / MACROS
  dind%dms_ind = 1; dind%dmsp_ind = 2; dind%no3_ind = 3; dind%doc_ind = 4; dind%zooC_ind = 5; dind%spC_ind = 6
  dind%spCaCO3_ind = 7; dind%diatC_ind = 8; dind%diazC_ind = 9; dind%phaeoC_ind = 10; dind%spChl_ind = 11
  dind%diatChl_ind = 12; dind%diazChl_ind = 13; dind%phaeoChl_ind = 14
  allocate(dind%short_name(DMS_tracer_cnt), dind%long_name(DMS_tracer_cnt), dind%units(DMS_tracer_cnt))
  call DMS_parms_init
  call DMS_init(dind)
  allocate(din%DMS_tracers(nL,nC,DMS_tracer_cnt), din%cell_thickness(nL,nC), din%number_of_active_levels(nC))
  read(iu) din%DMS_tracers
  din%cell_thickness = bin%cell_thickness
  din%number_of_active_levels = bin%number_of_active_levels
  allocate(dfo%ShortWaveFlux_surface(nC), dfo%surfacePressure(nC), dfo%iceFraction(nC), dfo%windSpeedSquared10m(nC), &
           dfo%SST(nC), dfo%SSS(nC), dfo%netFlux(nC,DMS_tracer_cnt))
  dfo%ShortWaveFlux_surface = bfo%ShortWaveFlux_surface; dfo%surfacePressure = bfo%surfacePressure
  dfo%iceFraction = bfo%iceFraction; dfo%windSpeedSquared10m = bfo%windSpeedSquared10m
  dfo%SST = bfo%SST; dfo%SSS = bfo%SSS; dfo%netFlux = 0.0_BGC_r8
  dfo%lcalc_DMS_gas_flux = .true.
  allocate(dout%DMS_tendencies(nL,nC,DMS_tracer_cnt)); dout%DMS_tendencies = fill_value

  mind%prot_ind = 1; mind%poly_ind = 2; mind%lip_ind = 3; mind%zooC_ind = 4; mind%spC_ind = 5
  mind%diatC_ind = 6; mind%diazC_ind = 7; mind%phaeoC_ind = 8
  allocate(mind%short_name(MACROS_tracer_cnt), mind%long_name(MACROS_tracer_cnt), mind%units(MACROS_tracer_cnt))
  call MACROS_parms_init
  call MACROS_init(mind)
  allocate(min_%MACROS_tracers(nL,nC,MACROS_tracer_cnt), min_%cell_thickness(nL,nC), min_%number_of_active_levels(nC))
  read(iu) min_%MACROS_tracers
  min_%cell_thickness = bin%cell_thickness
  min_%number_of_active_levels = bin%number_of_active_levels
  allocate(mout%MACROS_tendencies(nL,nC,MACROS_tracer_cnt)); mout%MACROS_tendencies = fill_value
  close(iu)

  ! ---- every diagnostic component, pre-filled with a sentinel (generated: gen_alloc.py)
  include 'alloc_diag.inc'

  ! ---- the reference itself: cold brackets, then warm brackets
  call BGC_SourceSink(autotrophs, bind, bin, bfo, bout, bdiag, nL, nC, nCols, alt)
  ph_cold = bout%PH_PREV_3D
  call BGC_SourceSink(autotrophs, bind, bin, bfo, bout, bdiag, nL, nC, nCols, alt)
  call BGC_SurfaceFluxes(bind, bin, bfo, bfdiag, nC, nCols)
  call DMS_SourceSink(dind, din, dfo, dout, ddiag, nL, nC, nCols)
  call DMS_SurfaceFluxes(dind, din, dfo, dfdiag, nC, nCols)
  call MACROS_SourceSink(mind, min_, mout, mdiag, nL, nC, nCols)

  open(ou, file=trim(fout), access='stream', form='unformatted', status='replace')
  write(ou) bout%BGC_tendencies, ph_cold, bout%PH_PREV_3D, bout%PH_PREV_ALT_CO2_3D
  write(ou) bfo%netFlux, bfo%gasFlux, bfo%surface_pH, bfo%surface_pH_alt_co2, bfo%iceFraction
  write(ou) dout%DMS_tendencies, dfo%netFlux, mout%MACROS_tendencies
  include 'write_diag.inc'
  close(ou)
end program ref_driver
