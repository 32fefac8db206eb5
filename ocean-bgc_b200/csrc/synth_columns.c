/* synth_columns.c — deterministic synthetic ocean columns (workload generator).
 *
 * Not part of the hot path and not the oracle: it only manufactures inputs of
 * the shapes BASELINE.json names (EC60to30 / RRS18to6) for tests and bench.py.
 * Values come from a counter-based hash keyed by (seed, field, GLOBAL column,
 * level), so CPU, 1-GPU and N-GPU shards see bit-identical inputs regardless of
 * how columns are sharded.  Profiles follow SURVEY.md section 8(d); units follow
 * BGC_mod.F90:323-328 (tracers mmol/m^3, depths cm, dust g/cm^2/s, SW W/m^2).
 */
#include "bgc_b200.h"
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct SynthSpec {
  uint64_t seed;
  int nLevelsMax, nColumnsMax, nColumns;
  long long column0;      /* global index of local column 0 (sharding)            */
  int nlev_active;        /* levels of a full-depth column (<= nLevelsMax)         */
  int ragged;             /* 0: every column has nlev_active levels; 1: ragged     */
  int soa;                /* 0: Fortran layout (k fastest); 1: SoA (column fastest) */
  int jitter;             /* 0: idealised profiles (config 1); 1: lognormal jitter  */
  int nthreads;
} SynthSpec;

static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
  x ^= x >> 27; x *= 0x94D049BB133111EBULL;
  x ^= x >> 31;
  return x;
}
static inline uint64_t key(uint64_t seed, uint64_t field, uint64_t col, uint64_t lev) {
  uint64_t x = mix64(seed + 0x9E3779B97F4A7C15ULL * (field + 1));
  x = mix64(x ^ (col + 0x632BE59BD9B4E019ULL));
  x = mix64(x ^ ((lev + 1) * 0xD6E8FEB86659FD93ULL));
  return x;
}
static inline double u01(uint64_t h) {   /* (0,1) */
  return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
static inline double nrm(uint64_t h) {   /* N(0,1), Box-Muller on two derived uniforms */
  double u1 = u01(h), u2 = u01(mix64(h ^ 0xA5A5A5A5A5A5A5A5ULL));
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

enum {   /* field ids for the hash */
  F_LAT = 1, F_KMAX, F_KMAX2, F_O2COL, F_SW, F_SW0, F_DUST, F_DUST0, F_U10, F_ICE, F_ICE2,
  F_PRES, F_NEG, F_NEGWHICH, F_TJ, F_SJ, F_DEP, F_RIV, F_ICEFLX,
  F_TRACER0 = 100, F_DMS0 = 200, F_MAC0 = 300
};

int bgc_synth_fill(const SynthSpec *sp, const BgcIndices *bi, const DmsIndices *di,
                   const MacrosIndices *mi, BgcInput *in, BgcForcing *fo, DmsInput *din,
                   DmsForcing *dfo, MacrosInput *min_) {
  const int nL = sp->nLevelsMax, nC = sp->nColumnsMax, nlev = sp->nlev_active;
  const int soa = sp->soa, jit = sp->jitter;
  const uint64_t seed = sp->seed;
  double dzv[512], zbotv[512], zmidv[512];
  double ssum = 0.0, a, acc = 0.0;
  int k, c;
  if (nlev > nL || nlev > 512 || nlev < 2) return BGC_ERR_ARG;

  /* vertical grid: 10 m surface cells stretching to 5500 m total */
  for (k = 0; k < nlev; ++k) { double s = (double)k / (double)(nlev - 1); ssum += s * s; }
  a = (550.0 - nlev) / ssum;
  for (k = 0; k < nlev; ++k) {
    double s = (double)k / (double)(nlev - 1);
    dzv[k] = 1000.0 * (1.0 + a * s * s);
    acc += dzv[k];
    zbotv[k] = acc;
    zmidv[k] = acc - 0.5 * dzv[k];
  }

#define IDX2(k_, c_) (soa ? ((size_t)(c_) + (size_t)nC * (size_t)(k_)) : ((size_t)(k_) + (size_t)nL * (size_t)(c_)))
#define IDX3(k_, c_, n_) (soa ? ((size_t)(c_) + (size_t)nC * ((size_t)(k_) + (size_t)nL * (size_t)(n_))) \
                              : ((size_t)(k_) + (size_t)nL * ((size_t)(c_) + (size_t)nC * (size_t)(n_))))
#define IDXF(c_, n_) ((size_t)(c_) + (size_t)nC * (size_t)(n_))

  {
    int nt = sp->nthreads;
#ifdef _OPENMP
    if (nt < 1) nt = omp_get_max_threads();
#else
    nt = 1;
#endif
#pragma omp parallel for num_threads(nt) schedule(static)
    for (c = 0; c < nC; ++c) {
      const uint64_t gc = (uint64_t)(sp->column0 + c);
      const int live = c < sp->nColumns;
      double lat_deg, lat, coslat, Ts, ucol;
      int kmax = nlev, kk, n;
      double tr[BGC_TRACER_CNT];

      lat_deg = -75.0 + 150.0 * u01(key(seed, F_LAT, gc, 0));
      if (!jit) lat_deg = 20.0;
      lat = lat_deg * (3.141592653589793 / 180.0);
      coslat = cos(lat);
      Ts = -1.8 + 31.0 * coslat * coslat;
      ucol = jit ? u01(key(seed, F_O2COL, gc, 0)) : 0.5;

      if (sp->ragged) {
        double r = u01(key(seed, F_KMAX, gc, 0));
        if (r < 0.005) kmax = 0;
        else if (r < 0.305) {
          int lo = nlev / 3;
          kmax = lo + (int)(u01(key(seed, F_KMAX2, gc, 0)) * (double)(nlev - lo));
          if (kmax > nlev - 1) kmax = nlev - 1;
        }
      }
      if (!live) kmax = 0;

      if (in) {
        if (in->number_of_active_levels) ((int *)in->number_of_active_levels)[c] = kmax;
        if (in->cell_latitude) ((double *)in->cell_latitude)[c] = lat;   /* radians; only the sign is used */
      }
      if (din && din->number_of_active_levels) ((int *)din->number_of_active_levels)[c] = kmax;
      if (min_ && min_->number_of_active_levels) ((int *)min_->number_of_active_levels)[c] = kmax;

      for (kk = 0; kk < nL; ++kk) {
        const int have = kk < nlev;
        const double zm = have ? zmidv[kk] * 0.01 : 0.0;   /* metres */
        double T, S, e800, nj[BGC_TRACER_CNT];
        double zC;

        if (!have) {   /* padding levels below the grid: benign values, never active */
          if (in) {
            if (in->PotentialTemperature) ((double *)in->PotentialTemperature)[IDX2(kk, c)] = 2.0;
            if (in->Salinity) ((double *)in->Salinity)[IDX2(kk, c)] = 34.7;
            if (in->cell_center_depth) ((double *)in->cell_center_depth)[IDX2(kk, c)] = 0.0;
            if (in->cell_thickness) ((double *)in->cell_thickness)[IDX2(kk, c)] = 1000.0;
            if (in->cell_bottom_depth) ((double *)in->cell_bottom_depth)[IDX2(kk, c)] = 0.0;
            if (in->BGC_tracers) for (n = 0; n < BGC_TRACER_CNT; ++n) ((double *)in->BGC_tracers)[IDX3(kk, c, n)] = 0.0;
          }
          if (fo && fo->FESEDFLUX) fo->FESEDFLUX[IDX2(kk, c)] = 0.0;
          if (din) {
            if (din->cell_thickness) ((double *)din->cell_thickness)[IDX2(kk, c)] = 1000.0;
            if (din->DMS_tracers) for (n = 0; n < DMS_TRACER_CNT; ++n) ((double *)din->DMS_tracers)[IDX3(kk, c, n)] = 0.0;
          }
          if (min_) {
            if (min_->cell_thickness) ((double *)min_->cell_thickness)[IDX2(kk, c)] = 1000.0;
            if (min_->MACROS_tracers) for (n = 0; n < MACROS_TRACER_CNT; ++n) ((double *)min_->MACROS_tracers)[IDX3(kk, c, n)] = 0.0;
          }
          continue;
        }

        e800 = exp(-zm / 800.0);
        T = Ts * e800 + 2.0 * (1.0 - e800);
        S = 34.7 + 0.8 * exp(-zm / 500.0) * cos(2.0 * lat);
        if (jit) {
          T += 0.3 * nrm(key(seed, F_TJ, gc, kk));
          S *= exp(0.005 * nrm(key(seed, F_SJ, gc, kk)));
          if (T < -1.9) T = -1.9;
        }

        for (n = 0; n < BGC_TRACER_CNT; ++n)
          nj[n] = jit ? exp(0.1 * nrm(key(seed, F_TRACER0 + n, gc, kk))) : 1.0;

        /* ---- BGC tracers (mmol/m^3), slot = host-chosen index - 1 */
        for (n = 0; n < BGC_TRACER_CNT; ++n) tr[n] = 0.0;
#define SET(ind_, val_) tr[(ind_)-1] = (val_) * nj[(ind_)-1]
        {
          double no3 = 0.05 + 35.0 * (1.0 - exp(-zm / 600.0));
          double doc = 4.0 + 60.0 * exp(-zm / 300.0);
          double dd = (zm - 800.0) / 500.0;
          double o2 = 280.0 - (200.0 + 80.0 * ucol) * exp(-dd * dd);
          double dic = 2000.0 + 250.0 * (1.0 - exp(-zm / 1000.0));
          double shal = (zm < 300.0) ? exp(-zm / 70.0) : 0.0;   /* exactly 0 below 300 m */
          double spC = 1.5 * shal, diatC = 0.8 * shal;
          double diazC = (Ts > 15.0) ? 0.05 * shal : 0.0;
          double phaeoC = (Ts < 8.0) ? 0.3 * shal : 0.0;
          if (o2 < 0.0) o2 = 0.0;
          SET(bi->no3_ind, no3);
          SET(bi->po4_ind, no3 / 16.0 + 0.02);
          SET(bi->sio3_ind, 1.0 + 120.0 * (1.0 - exp(-zm / 1500.0)));
          SET(bi->nh4_ind, 0.3 * exp(-zm / 100.0));
          SET(bi->fe_ind, 1e-4 + 6e-4 * (1.0 - exp(-zm / 500.0)));
          SET(bi->o2_ind, o2);
          SET(bi->dic_ind, dic);
          SET(bi->dic_alt_co2_ind, dic - 15.0);
          SET(bi->alk_ind, 2300.0 + 100.0 * (1.0 - exp(-zm / 2000.0)));
          SET(bi->doc_ind, doc);
          SET(bi->don_ind, doc * 0.07);
          SET(bi->dop_ind, doc * 0.004);
          SET(bi->dofe_ind, doc * 2e-6);
          SET(bi->donr_ind, 1.5);
          SET(bi->dopr_ind, 0.03);
          SET(bi->zooC_ind, 1.0 * exp(-zm / 80.0));
          SET(bi->spC_ind, spC);
          SET(bi->spChl_ind, 0.25 * spC);
          SET(bi->spFe_ind, 6e-6 * spC);
          SET(bi->spCaCO3_ind, 0.05 * spC);
          SET(bi->diatC_ind, diatC);
          SET(bi->diatChl_ind, 0.25 * diatC);
          SET(bi->diatFe_ind, 6e-6 * diatC);
          SET(bi->diatSi_ind, 0.137 * diatC);
          SET(bi->diazC_ind, diazC);
          SET(bi->diazChl_ind, 0.25 * diazC);
          SET(bi->diazFe_ind, 6e-6 * diazC);
          SET(bi->phaeoC_ind, phaeoC);
          SET(bi->phaeoChl_ind, 0.25 * phaeoC);
          SET(bi->phaeoFe_ind, 6e-6 * phaeoC);
          /* DIC/ALK carry a much smaller jitter so the carbonate system stays oceanic */
          if (jit) {
            tr[bi->dic_ind - 1] = dic * (1.0 + 0.01 * (nj[bi->dic_ind - 1] - 1.0));
            tr[bi->dic_alt_co2_ind - 1] = (dic - 15.0) * (1.0 + 0.01 * (nj[bi->dic_alt_co2_ind - 1] - 1.0));
            tr[bi->alk_ind - 1] = (2300.0 + 100.0 * (1.0 - exp(-zm / 2000.0))) *
                                  (1.0 + 0.01 * (nj[bi->alk_ind - 1] - 1.0));
          }
        }
#undef SET
        /* 1 % of cells: one tracer slightly negative (exercises max(0, .)) */
        if (jit && u01(key(seed, F_NEG, gc, kk)) < 0.01) {
          int which = (int)(u01(key(seed, F_NEGWHICH, gc, kk)) * BGC_TRACER_CNT);
          if (which >= BGC_TRACER_CNT) which = BGC_TRACER_CNT - 1;
          tr[which] = -1e-3;
        }
        zC = tr[bi->zooC_ind - 1];

        if (in) {
          if (in->PotentialTemperature) ((double *)in->PotentialTemperature)[IDX2(kk, c)] = T;
          if (in->Salinity) ((double *)in->Salinity)[IDX2(kk, c)] = S;
          if (in->cell_center_depth) ((double *)in->cell_center_depth)[IDX2(kk, c)] = zmidv[kk];
          if (in->cell_thickness) ((double *)in->cell_thickness)[IDX2(kk, c)] = dzv[kk];
          if (in->cell_bottom_depth) ((double *)in->cell_bottom_depth)[IDX2(kk, c)] = zbotv[kk];
          if (in->BGC_tracers)
            for (n = 0; n < BGC_TRACER_CNT; ++n) ((double *)in->BGC_tracers)[IDX3(kk, c, n)] = tr[n];
        }
        if (fo && fo->FESEDFLUX) fo->FESEDFLUX[IDX2(kk, c)] = (kk == kmax - 1) ? 2.3e-6 : 0.0;

        if (din) {
          if (din->cell_thickness) ((double *)din->cell_thickness)[IDX2(kk, c)] = dzv[kk];
          if (din->DMS_tracers) {
            double *t = (double *)din->DMS_tracers;
            double j1 = jit ? exp(0.1 * nrm(key(seed, F_DMS0 + 0, gc, kk))) : 1.0;
            double j2 = jit ? exp(0.1 * nrm(key(seed, F_DMS0 + 1, gc, kk))) : 1.0;
            for (n = 0; n < DMS_TRACER_CNT; ++n) t[IDX3(kk, c, n)] = 0.0;
            t[IDX3(kk, c, di->dms_ind - 1)] = 2e-3 * exp(-zm / 100.0) * j1;
            t[IDX3(kk, c, di->dmsp_ind - 1)] = 5e-3 * exp(-zm / 100.0) * j2;
            t[IDX3(kk, c, di->no3_ind - 1)] = tr[bi->no3_ind - 1];
            t[IDX3(kk, c, di->doc_ind - 1)] = tr[bi->doc_ind - 1];
            t[IDX3(kk, c, di->zooC_ind - 1)] = zC;
            t[IDX3(kk, c, di->spC_ind - 1)] = tr[bi->spC_ind - 1];
            t[IDX3(kk, c, di->spCaCO3_ind - 1)] = tr[bi->spCaCO3_ind - 1];
            t[IDX3(kk, c, di->diatC_ind - 1)] = tr[bi->diatC_ind - 1];
            t[IDX3(kk, c, di->diazC_ind - 1)] = tr[bi->diazC_ind - 1];
            t[IDX3(kk, c, di->phaeoC_ind - 1)] = tr[bi->phaeoC_ind - 1];
            t[IDX3(kk, c, di->spChl_ind - 1)] = tr[bi->spChl_ind - 1];
            t[IDX3(kk, c, di->diatChl_ind - 1)] = tr[bi->diatChl_ind - 1];
            t[IDX3(kk, c, di->diazChl_ind - 1)] = tr[bi->diazChl_ind - 1];
            t[IDX3(kk, c, di->phaeoChl_ind - 1)] = tr[bi->phaeoChl_ind - 1];
          }
        }
        if (min_) {
          if (min_->cell_thickness) ((double *)min_->cell_thickness)[IDX2(kk, c)] = dzv[kk];
          if (min_->MACROS_tracers) {
            double *t = (double *)min_->MACROS_tracers;
            double e200 = exp(-zm / 200.0);
            double j1 = jit ? exp(0.1 * nrm(key(seed, F_MAC0 + 0, gc, kk))) : 1.0;
            double j2 = jit ? exp(0.1 * nrm(key(seed, F_MAC0 + 1, gc, kk))) : 1.0;
            double j3 = jit ? exp(0.1 * nrm(key(seed, F_MAC0 + 2, gc, kk))) : 1.0;
            t[IDX3(kk, c, mi->prot_ind - 1)] = 1.0 * e200 * j1;
            t[IDX3(kk, c, mi->poly_ind - 1)] = 3.0 * e200 * j2;
            t[IDX3(kk, c, mi->lip_ind - 1)] = 0.3 * e200 * j3;
            t[IDX3(kk, c, mi->zooC_ind - 1)] = zC;
            t[IDX3(kk, c, mi->spC_ind - 1)] = tr[bi->spC_ind - 1];
            t[IDX3(kk, c, mi->diatC_ind - 1)] = tr[bi->diatC_ind - 1];
            t[IDX3(kk, c, mi->diazC_ind - 1)] = tr[bi->diazC_ind - 1];
            t[IDX3(kk, c, mi->phaeoC_ind - 1)] = tr[bi->phaeoC_ind - 1];
          }
        }
      }   /* levels */

      /* ---- per-column forcing */
      {
        double sw = 350.0 * coslat * (jit ? u01(key(seed, F_SW, gc, 0)) : 0.7);
        double dustv, u10, ice, pres;
        double sst, sss;
        if (jit && u01(key(seed, F_SW0, gc, 0)) < 0.20) sw = 0.0;
        dustv = 1e-11 * (jit ? exp(1.0 * nrm(key(seed, F_DUST, gc, 0))) : 1.0);
        if (jit && u01(key(seed, F_DUST0, gc, 0)) < 0.05) dustv = 0.0;
        u10 = jit ? 1.0 + (2.25e6 - 1.0) * u01(key(seed, F_U10, gc, 0)) : 0.5e6;
        ice = 0.0;
        if (jit && u01(key(seed, F_ICE, gc, 0)) >= 0.80) ice = u01(key(seed, F_ICE2, gc, 0));
        pres = jit ? 1.0 + 0.03 * (2.0 * u01(key(seed, F_PRES, gc, 0)) - 1.0) : 1.0;
        {   /* SST/SSS = level-1 T/S (recomputed exactly as above) */
          double zm0 = zmidv[0] * 0.01, e0 = exp(-zm0 / 800.0);
          sst = Ts * e0 + 2.0 * (1.0 - e0);
          sss = 34.7 + 0.8 * exp(-zm0 / 500.0) * cos(2.0 * lat);
          if (jit) {
            sst += 0.3 * nrm(key(seed, F_TJ, gc, 0));
            sss *= exp(0.005 * nrm(key(seed, F_SJ, gc, 0)));
            if (sst < -1.9) sst = -1.9;
          }
        }
        if (fo) {
          if (fo->dust_FLUX_IN) fo->dust_FLUX_IN[c] = dustv;
          if (fo->ShortWaveFlux_surface) fo->ShortWaveFlux_surface[c] = sw;
          if (fo->surfacePressure) fo->surfacePressure[c] = pres;
          if (fo->iceFraction) fo->iceFraction[c] = ice;
          if (fo->windSpeedSquared10m) fo->windSpeedSquared10m[c] = u10;
          if (fo->atmCO2) fo->atmCO2[c] = 400.0;
          if (fo->atmCO2_ALT_CO2) fo->atmCO2_ALT_CO2[c] = 284.7;
          if (fo->surface_pH) fo->surface_pH[c] = 0.0;
          if (fo->surface_pH_alt_co2) fo->surface_pH_alt_co2[c] = 0.0;
          if (fo->surfaceDepth) fo->surfaceDepth[c] = 5.0;
          if (fo->SST) fo->SST[c] = sst;
          if (fo->SSS) fo->SSS[c] = sss;
          for (n = 0; n < BGC_TRACER_CNT; ++n) {
            if (fo->depositionFlux) fo->depositionFlux[IDXF(c, n)] = jit ? 1e-9 * u01(key(seed, F_DEP, gc, n)) : 1e-9;
            if (fo->riverFlux) fo->riverFlux[IDXF(c, n)] = jit ? 1e-10 * u01(key(seed, F_RIV, gc, n)) : 1e-10;
            if (fo->seaIceFlux) fo->seaIceFlux[IDXF(c, n)] = jit ? 1e-11 * (2.0 * u01(key(seed, F_ICEFLX, gc, n)) - 1.0) : 0.0;
            if (fo->gasFlux) fo->gasFlux[IDXF(c, n)] = 0.0;
            if (fo->netFlux) fo->netFlux[IDXF(c, n)] = 0.0;
          }
        }
        if (dfo) {
          if (dfo->ShortWaveFlux_surface) dfo->ShortWaveFlux_surface[c] = sw;
          if (dfo->surfacePressure) dfo->surfacePressure[c] = pres;
          if (dfo->iceFraction) dfo->iceFraction[c] = ice;
          if (dfo->windSpeedSquared10m) dfo->windSpeedSquared10m[c] = u10;
          if (dfo->SST) dfo->SST[c] = sst;
          if (dfo->SSS) dfo->SSS[c] = sss;
          if (dfo->netFlux) for (n = 0; n < DMS_TRACER_CNT; ++n) dfo->netFlux[IDXF(c, n)] = 0.0;
        }
      }
    }   /* columns */
  }
  return BGC_OK;
}

/* Config 2: n surface points for co2calc_1point.  out[11][n]:
 * depth,temp,salt,dic,ta,pt,sit,phlo,phhi,xco2,atmpres */
int bgc_synth_co2_points(uint64_t seed, long long i0, int n, double *out) {
  int i;
#pragma omp parallel for schedule(static)
  for (i = 0; i < n; ++i) {
    uint64_t g = (uint64_t)(i0 + i);
    double dic = 1800.0 + 500.0 * u01(key(seed, 1, g, 0));
    out[0 * (size_t)n + i] = 5.0;
    out[1 * (size_t)n + i] = -1.8 + 32.8 * u01(key(seed, 2, g, 0));
    out[2 * (size_t)n + i] = 30.0 + 8.0 * u01(key(seed, 3, g, 0));
    out[3 * (size_t)n + i] = dic;
    out[4 * (size_t)n + i] = dic + 80.0 + 340.0 * u01(key(seed, 4, g, 0));
    out[5 * (size_t)n + i] = 3.0 * u01(key(seed, 5, g, 0));
    out[6 * (size_t)n + i] = 150.0 * u01(key(seed, 6, g, 0));
    out[7 * (size_t)n + i] = 7.0;
    out[8 * (size_t)n + i] = 9.0;
    out[9 * (size_t)n + i] = 280.0 + 280.0 * u01(key(seed, 7, g, 0));
    out[10 * (size_t)n + i] = 0.95 + 0.10 * u01(key(seed, 8, g, 0));
  }
  return BGC_OK;
}
