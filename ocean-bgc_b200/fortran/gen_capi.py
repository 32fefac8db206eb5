#!/usr/bin/env python
"""Generates the ISO_C_BINDING mirror of include/bgc_b200.h:

    bgc_b200_capi.F90      bind(C) derived types + interfaces of every C-ABI entry point
    bgc_diag_ptrs.inc      cdg%<member> = loc2/loc3/loc1(d%<member>)  for BGC_diagnostics_type
    dms_diag_ptrs.inc, macros_diag_ptrs.inc, bgc_flux_diag_ptrs.inc, dms_flux_diag_ptrs.inc

The header is the single source of truth (the Python ctypes mirror, ocean-bgc_b200/abi.py,
parses the same file), so the Fortran shim cannot drift from the C ABI.  Run after any
change to the header:   python ocean-bgc_b200/fortran/gen_capi.py
"""
import importlib.util
import os
import ctypes as C

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
spec = importlib.util.spec_from_file_location("abi", os.path.join(PKG, "abi.py"))
abi = importlib.util.module_from_spec(spec)
spec.loader.exec_module(abi)

FT = {C.c_double: "real(c_double)", C.c_int: "integer(c_int)", C.c_ulonglong: "integer(c_long_long)"}


def ftype(name, fields):
    out = ["  type, bind(C) :: %s" % name]
    for fname, typ in fields:
        if hasattr(typ, "_length_"):                      # fixed-size array
            out.append("    %s :: %s(%d)" % (FT[typ._type_], fname, typ._length_))
        elif hasattr(typ, "contents") or typ.__name__.startswith("LP_"):
            out.append("    type(c_ptr) :: %s = c_null_ptr" % fname)
        else:
            out.append("    %s :: %s" % (FT[typ], fname))
    out.append("  end type %s" % name)
    return "\n".join(out)


STRUCTS = ["BgcParams", "BgcAutotroph", "BgcIndices", "DmsParams", "DmsIndices", "MacrosParams",
           "MacrosIndices", "BgcInput", "BgcForcing", "BgcOutput", "BgcFluxDiagnostics",
           "BgcDiagnostics", "DmsInput", "DmsForcing", "DmsOutput", "DmsFluxDiagnostics",
           "DmsDiagnostics", "MacrosInput", "MacrosOutput", "MacrosDiagnostics", "BgcStatus"]

IFACES = r'''
  interface
    integer(c_int) function bgc_ctx_create(device, nLevelsMax, nColumnsMax, ctx) bind(C, name="bgc_ctx_create")
      import :: c_int, c_ptr
      integer(c_int), value :: device, nLevelsMax, nColumnsMax
      type(c_ptr), intent(out) :: ctx
    end function
    integer(c_int) function bgc_ctx_destroy(ctx) bind(C, name="bgc_ctx_destroy")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
    end function
    integer(c_int) function bgc_ctx_synchronize(ctx) bind(C, name="bgc_ctx_synchronize")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
    end function
    integer(c_int) function bgc_get_status(ctx, st, reset) bind(C, name="bgc_get_status")
      import :: c_int, c_ptr, BgcStatus
      type(c_ptr), value :: ctx
      type(BgcStatus), intent(out) :: st
      integer(c_int), value :: reset
    end function
    function bgc_last_error() bind(C, name="bgc_last_error") result(msg)
      import :: c_ptr
      type(c_ptr) :: msg
    end function
    integer(c_int) function bgc_set_params(ctx, p, autotrophs, ind) bind(C, name="bgc_set_params")
      import :: c_int, c_ptr, BgcParams, BgcAutotroph, BgcIndices
      type(c_ptr), value :: ctx
      type(BgcParams), intent(in) :: p
      type(BgcAutotroph), intent(in) :: autotrophs(4)
      type(BgcIndices), intent(in) :: ind
    end function
    integer(c_int) function dms_set_params(ctx, p, ind) bind(C, name="dms_set_params")
      import :: c_int, c_ptr, DmsParams, DmsIndices
      type(c_ptr), value :: ctx
      type(DmsParams), intent(in) :: p
      type(DmsIndices), intent(in) :: ind
    end function
    integer(c_int) function macros_set_params(ctx, p, ind) bind(C, name="macros_set_params")
      import :: c_int, c_ptr, MacrosParams, MacrosIndices
      type(c_ptr), value :: ctx
      type(MacrosParams), intent(in) :: p
      type(MacrosIndices), intent(in) :: ind
    end function
    integer(c_int) function bgc_source_sink(ctx, cin, cfo, cout, cdg, nLevelsMax, nColumnsMax, nColumns, &
                                            alt_co2_use_eco, mem_space) bind(C, name="bgc_source_sink")
      import :: c_int, c_ptr, BgcInput, BgcForcing, BgcOutput, BgcDiagnostics
      type(c_ptr), value :: ctx
      type(BgcInput), intent(in) :: cin
      type(BgcForcing), intent(in) :: cfo
      type(BgcOutput), intent(inout) :: cout
      type(BgcDiagnostics), intent(inout) :: cdg
      integer(c_int), value :: nLevelsMax, nColumnsMax, nColumns, alt_co2_use_eco, mem_space
    end function
    integer(c_int) function bgc_surface_fluxes(ctx, cin, cfo, cfd, nLevelsMax, nColumnsMax, nColumns, mem_space) &
        bind(C, name="bgc_surface_fluxes")
      import :: c_int, c_ptr, BgcInput, BgcForcing, BgcFluxDiagnostics
      type(c_ptr), value :: ctx
      type(BgcInput), intent(in) :: cin
      type(BgcForcing), intent(inout) :: cfo
      type(BgcFluxDiagnostics), intent(inout) :: cfd
      integer(c_int), value :: nLevelsMax, nColumnsMax, nColumns, mem_space
    end function
    integer(c_int) function dms_source_sink(ctx, cin, cfo, cout, cdg, nLevelsMax, nColumnsMax, nColumns, mem_space) &
        bind(C, name="dms_source_sink")
      import :: c_int, c_ptr, DmsInput, DmsForcing, DmsOutput, DmsDiagnostics
      type(c_ptr), value :: ctx
      type(DmsInput), intent(in) :: cin
      type(DmsForcing), intent(in) :: cfo
      type(DmsOutput), intent(inout) :: cout
      type(DmsDiagnostics), intent(inout) :: cdg
      integer(c_int), value :: nLevelsMax, nColumnsMax, nColumns, mem_space
    end function
    integer(c_int) function dms_surface_fluxes(ctx, cin, cfo, cfd, nLevelsMax, nColumnsMax, nColumns, mem_space) &
        bind(C, name="dms_surface_fluxes")
      import :: c_int, c_ptr, DmsInput, DmsForcing, DmsFluxDiagnostics
      type(c_ptr), value :: ctx
      type(DmsInput), intent(in) :: cin
      type(DmsForcing), intent(inout) :: cfo
      type(DmsFluxDiagnostics), intent(inout) :: cfd
      integer(c_int), value :: nLevelsMax, nColumnsMax, nColumns, mem_space
    end function
    integer(c_int) function macros_source_sink(ctx, cin, cout, cdg, nLevelsMax, nColumnsMax, nColumns, mem_space) &
        bind(C, name="macros_source_sink")
      import :: c_int, c_ptr, MacrosInput, MacrosOutput, MacrosDiagnostics
      type(c_ptr), value :: ctx
      type(MacrosInput), intent(in) :: cin
      type(MacrosOutput), intent(inout) :: cout
      type(MacrosDiagnostics), intent(inout) :: cdg
      integer(c_int), value :: nLevelsMax, nColumnsMax, nColumns, mem_space
    end function
    integer(c_int) function bgc_co2calc_points(ctx, n, depth, temp, salt, dic, ta, pt, sit, phlo, phhi, xco2, &
                                               atmpres, ph, co2star, dco2star, pco2surf, dpco2, mem_space) &
        bind(C, name="bgc_co2calc_points")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: n, mem_space
      real(c_double), intent(in) :: depth(*), temp(*), salt(*), dic(*), ta(*), pt(*), sit(*), phlo(*), phhi(*), &
                                    xco2(*), atmpres(*)
      real(c_double), intent(inout) :: ph(*), co2star(*), dco2star(*), pco2surf(*), dpco2(*)
    end function
    integer(c_int) function bgc_comp_co3terms(ctx, n, k_level, k_all, depth, temp, salt, dic, ta, pt, sit, phlo, phhi, &
                                              ph, h2co3, hco3, co3, mem_space) bind(C, name="bgc_comp_co3terms")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: n, k_all, mem_space
      integer(c_int), intent(in) :: k_level(*)
      real(c_double), intent(in) :: depth(*), temp(*), salt(*), dic(*), ta(*), pt(*), sit(*), phlo(*), phhi(*)
      real(c_double), intent(inout) :: ph(*), h2co3(*), hco3(*), co3(*)
    end function
    integer(c_int) function bgc_comp_co3_sat_vals(ctx, n, k_level, k_all, depth, temp, salt, co3_sat_calc, &
                                                  co3_sat_arag, mem_space) bind(C, name="bgc_comp_co3_sat_vals")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      integer(c_int), value :: n, k_all, mem_space
      integer(c_int), intent(in) :: k_level(*)
      real(c_double), intent(in) :: depth(*), temp(*), salt(*)
      real(c_double), intent(inout) :: co3_sat_calc(*), co3_sat_arag(*)
    end function
    integer(c_int) function bgc_state_set(ctx, which, host, nLevelsMax, nColumnsMax) bind(C, name="bgc_state_set")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, host
      integer(c_int), value :: which, nLevelsMax, nColumnsMax
    end function
    integer(c_int) function bgc_state_get(ctx, which, host, nLevelsMax, nColumnsMax) bind(C, name="bgc_state_get")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx, host
      integer(c_int), value :: which, nLevelsMax, nColumnsMax
    end function
    integer(c_int) function bgc_state_device_ptr(ctx, which, nLevelsMax, nColumnsMax, dev_ptr) &
        bind(C, name="bgc_state_device_ptr")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: which, nLevelsMax, nColumnsMax
      type(c_ptr), intent(out) :: dev_ptr
    end function
    integer(c_int) function bgc_inventory_enable(ctx, enable) bind(C, name="bgc_inventory_enable")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
      integer(c_int), value :: enable
    end function
    integer(c_int) function bgc_inventory_reset(ctx) bind(C, name="bgc_inventory_reset")
      import :: c_int, c_ptr
      type(c_ptr), value :: ctx
    end function
    integer(c_int) function bgc_inventory_allreduce(ctx, vec) bind(C, name="bgc_inventory_allreduce")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: ctx
      real(c_double), intent(out) :: vec(64)
    end function
    integer(c_int) function bgc_comm_unique_id(id) bind(C, name="bgc_comm_unique_id")
      import :: c_int, c_signed_char
      integer(c_signed_char), intent(out) :: id(128)
    end function
    integer(c_int) function bgc_comm_init_rank(ctx, nranks, rank, id) bind(C, name="bgc_comm_init_rank")
      import :: c_int, c_ptr, c_signed_char
      type(c_ptr), value :: ctx
      integer(c_int), value :: nranks, rank
      integer(c_signed_char), intent(in) :: id(128)
    end function
    integer(c_int) function bgc_host_register(ptr, bytes) bind(C, name="bgc_host_register")
      import :: c_int, c_ptr, c_size_t
      type(c_ptr), value :: ptr
      integer(c_size_t), value :: bytes
    end function
  end interface
'''


def main():
    out = ["! GENERATED by gen_capi.py from include/bgc_b200.h - do not edit by hand.",
           "! ISO_C_BINDING mirror of the C ABI of libbgc_b200.so (see INTEGRATION.md).",
           "module bgc_b200_capi",
           "  use, intrinsic :: iso_c_binding",
           "  implicit none",
           "  public",
           "  integer(c_int), parameter :: BGC_MEM_HOST_FORTRAN = %d, BGC_MEM_DEVICE_SOA = %d" %
           (abi.BGC_MEM_HOST_FORTRAN, abi.BGC_MEM_DEVICE_SOA),
           "  integer(c_int), parameter :: BGC_OK = 0, BGC_INVENTORY_LEN = %d" % abi.BGC_INVENTORY_LEN,
           "  integer(c_int), parameter :: BGC_STATE_PH_PREV_3D = 0, BGC_STATE_PH_PREV_ALT_CO2_3D = 1, &",
           "                               BGC_STATE_SURFACE_PH = 2, BGC_STATE_SURFACE_PH_ALT_CO2 = 3", ""]
    for s in STRUCTS:
        out.append(ftype(s, abi._STRUCT_FIELDS[s]))
        out.append("")
    out.append(IFACES)
    out.append("end module bgc_b200_capi")
    open(os.path.join(HERE, "bgc_b200_capi.F90"), "w").write("\n".join(out) + "\n")

    def inc(fname, lists, dst, src):
        lines = ["! GENERATED by gen_capi.py: C pointers to the allocatable components (NULL when not allocated)"]
        for names, fn in lists:
            for n in names:
                lines.append("  %s%%%s = %s(%s%%%s)" % (dst, n, fn, src, n))
        open(os.path.join(HERE, fname), "w").write("\n".join(lines) + "\n")

    inc("bgc_diag_ptrs.inc", [(abi.BGC_DIAG_K2, "loc2"), (abi.BGC_DIAG_KA, "loc3"), (abi.BGC_DIAG_CA, "loc2"),
                              (abi.BGC_DIAG_C1, "loc1")], "cdg", "BGC_diagnostic_fields")
    inc("bgc_flux_diag_ptrs.inc", [(abi.BGC_FLUX_DIAG, "loc1")], "cfd", "BGC_flux_diagnostic_fields")
    inc("dms_diag_ptrs.inc", [(abi.DMS_DIAG, "loc2")], "cdg", "DMS_diagnostic_fields")
    inc("dms_flux_diag_ptrs.inc", [(abi.DMS_FLUX_DIAG, "loc1")], "cfd", "DMS_flux_diagnostic_fields")
    inc("macros_diag_ptrs.inc", [(abi.MACROS_DIAG, "loc2")], "cdg", "MACROS_diagnostic_fields")
    print("generated bgc_b200_capi.F90 and 5 include files in", HERE)


if __name__ == "__main__":
    main()
