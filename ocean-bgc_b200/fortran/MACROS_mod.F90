!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
! MACROS_mod - drop-in replacement of the reference module of the same name
! (MACROS_mod.F90): same public entities and signatures, MACROS_parms unchanged.
!
!   MACROS_SourceSink  (ref. MACROS_mod.F90:137-411) -> macros_source_sink
!   MACROS_init        (ref. MACROS_mod.F90:72-129)  host-side metadata
!|||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||||
module MACROS_mod
  use, intrinsic :: iso_c_binding
  use MACROS_parms
  use bgc_b200_capi
  use bgc_b200_runtime
  implicit none
  private

  public :: MACROS_tracer_cnt, MACROS_init, MACROS_SourceSink

  integer (MACROS_i4), parameter :: MACROS_tracer_cnt = 8

contains

  subroutine MACROS_init(MACROS_indices)
    type(MACROS_indices_type), intent(inout) :: MACROS_indices
    call meta(MACROS_indices%prot_ind,   'PROT',   'Proteins')
    call meta(MACROS_indices%poly_ind,   'POLY',   'Polysaccharides')
    call meta(MACROS_indices%lip_ind,    'LIP',    'Lipids')
    call meta(MACROS_indices%zooC_ind,   'zooC',   'Zooplankton Carbon')
    call meta(MACROS_indices%spC_ind,    'spC',    ' Small Phytoplankton Carbon')
    call meta(MACROS_indices%diatC_ind,  'diatC',  ' Diatom Carbon')
    call meta(MACROS_indices%diazC_ind,  'diazC',  ' Diazotroph Carbon')
    call meta(MACROS_indices%phaeoC_ind, 'phaeoC', 'Phaeocystis Carbon')
    MACROS_indices%units(:) = 'mmol/m^3'
  contains
    subroutine meta(ind, sname, lname)
      integer (MACROS_i4), intent(in) :: ind
      character(len=*), intent(in) :: sname, lname
      MACROS_indices%short_name(ind) = sname
      MACROS_indices%long_name(ind) = lname
    end subroutine meta
  end subroutine MACROS_init

  subroutine MACROS_SourceSink(MACROS_indices, MACROS_input, MACROS_output, MACROS_diagnostic_fields, &
                               numLevelsMax, numColumnsMax, numColumns)
    type(MACROS_indices_type),     intent(in )           :: MACROS_indices
    type(MACROS_input_type),       intent(in ), target   :: MACROS_input
    integer (MACROS_i4), intent(in) :: numLevelsMax, numColumnsMax, numColumns
    type(MACROS_output_type),      intent(inout), target :: MACROS_output
    type(MACROS_diagnostics_type), intent(inout), target :: MACROS_diagnostic_fields
    type(c_ptr) :: ctx
    logical :: fresh
    type(MacrosParams) :: p
    type(MacrosIndices) :: ci
    type(MacrosInput) :: cin
    type(MacrosOutput) :: cout
    type(MacrosDiagnostics) :: cdg

    ctx = bgc_b200_ctx(numLevelsMax, numColumnsMax, fresh)
    p%f_prot = f_prot;  p%f_poly = f_poly;  p%f_lip = f_lip;  p%k_C_p_base = k_C_p_base
    p%zooC_avg = zooC_avg;  p%mort = mort;  p%k_prot_bac = k_prot_bac;  p%k_poly_bac = k_poly_bac
    p%k_lip_bac = k_lip_bac;  p%inject_scale = inject_scale
    ci%prot_ind = MACROS_indices%prot_ind;  ci%poly_ind = MACROS_indices%poly_ind
    ci%lip_ind = MACROS_indices%lip_ind;    ci%zooC_ind = MACROS_indices%zooC_ind
    ci%spC_ind = MACROS_indices%spC_ind;    ci%diatC_ind = MACROS_indices%diatC_ind
    ci%diazC_ind = MACROS_indices%diazC_ind; ci%phaeoC_ind = MACROS_indices%phaeoC_ind
    call bgc_b200_check(macros_set_params(ctx, p, ci), 'macros_set_params')

    cin%MACROS_tracers = loc3(MACROS_input%MACROS_tracers)
    cin%cell_thickness = loc2(MACROS_input%cell_thickness)
    cin%number_of_active_levels = loci1(MACROS_input%number_of_active_levels)
    cout%MACROS_tendencies = loc3(MACROS_output%MACROS_tendencies)
    include 'macros_diag_ptrs.inc'
    call bgc_b200_check(macros_source_sink(ctx, cin, cout, cdg, int(numLevelsMax, c_int),          &
                                           int(numColumnsMax, c_int), int(numColumns, c_int),      &
                                           BGC_MEM_HOST_FORTRAN), 'macros_source_sink')
  end subroutine MACROS_SourceSink

end module MACROS_mod
