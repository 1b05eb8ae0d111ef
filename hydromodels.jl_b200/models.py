"""Model definitions for the BASELINE configs, written with the component mirror.

Equations are transcribed from the reference's own model scripts:

* ExpHydro      — `/root/reference/test/models/exphydro.jl:8-37` (README variant) and
                  `/root/reference/docs/src/get_start_en.md:105-134` (`pet` declared first)
* GR4J + 2 UH   — `/root/reference/test/base/run_single_lumped.jl:106-156`
* HBV           — dPL-HBV tutorial `/root/reference/docs/src/tutorials/neuralnetwork_embeding.md:160-201`
                  with `BETA` as a plain parameter (HBV-edu itself lives in the external
                  HydroModelLibrary; parity for it is oracle-pinned)
* M50           — `/root/reference/test/base/run_single_lumped.jl:178-251`
* grid ExpHydro + D8 route — `/root/reference/test/base/run_spatial_model.jl:10-67`
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from .components import (Chain, Dense, HydroBucket, HydroFlux, HydroModel, HydroRoute, NeuralFlux,
                         StateFlux, UnitHydrograph, build_aggr_func)
from .symbolic import Symbols

PET_HAMON = "pet ~ 29.8 * lday * 24 * 0.611 * exp((17.3 * temp) / (temp + 237.3)) / (temp + 273.2)"


def _exphydro_buckets(htypes, docs_variant: bool):
    s = Symbols("temp lday pet prcp snowfall rainfall snowpack melt soilwater evap baseflow surfaceflow flow",
                "Tmin Tmax Df Smax Qmax f")
    split = HydroFlux(["snowfall ~ step_func(Tmin - temp) * prcp",
                       "rainfall ~ step_func(temp - Tmin) * prcp"], symbols=s)
    melt = HydroFlux("melt ~ step_func(temp - Tmax) * step_func(snowpack) * min(snowpack, Df * (temp - Tmax))",
                     symbols=s)
    pet = HydroFlux(PET_HAMON, symbols=s)
    snow_fluxes = [pet, split, melt] if docs_variant else [split, melt, pet]
    bucket_1 = HydroBucket(name="surface", fluxes=snow_fluxes,
                           dfluxes=[StateFlux("snowpack ~ snowfall - melt", symbols=s)], htypes=htypes)
    bucket_2 = HydroBucket(
        name="soil",
        fluxes=[HydroFlux("evap ~ step_func(soilwater) * pet * min(1.0, soilwater / Smax)", symbols=s),
                HydroFlux("baseflow ~ step_func(soilwater) * Qmax * exp(-f * (max(0.0, Smax - soilwater)))",
                          symbols=s),
                HydroFlux("surfaceflow ~ max(0.0, soilwater - Smax)", symbols=s),
                HydroFlux("flow ~ baseflow + surfaceflow", symbols=s)],
        dfluxes=[StateFlux("soilwater ~ (rainfall + melt) - (evap + flow)", symbols=s)], htypes=htypes)
    return s, bucket_1, bucket_2


def exphydro(htypes=None, docs_variant: bool = False) -> HydroModel:
    """ExpHydro, two buckets.  Inputs `temp, lday, prcp` (row order of the reference's test
    loader, `test/test_helpers.jl:27`); params `Tmin Tmax Df Smax Qmax f`; rows
    `snowpack soilwater | snowfall rainfall melt pet evap baseflow surfaceflow flow`
    (`docs_variant=True`: `pet` first, the order of the published table,
    `docs/src/get_start_en.md:379`)."""
    _, b1, b2 = _exphydro_buckets(htypes, docs_variant)
    return HydroModel(name="exphydro", components=(b1, b2), input_names=["temp", "lday", "prcp"])


def exphydro_single_bucket(htypes=None) -> HydroModel:
    """The one-bucket/two-state variant (`test/models/exphydrov2.jl:8-27`): numerically
    *different* from the two-bucket model — used to test that the engine
    honours component boundaries."""
    s = Symbols("temp lday pet prcp snowfall rainfall snowpack melt soilwater evap baseflow surfaceflow flow",
                "Tmin Tmax Df Smax Qmax f")
    b = HydroBucket(
        name="exphydro_v2",
        fluxes=[HydroFlux(["snowfall ~ step_func(Tmin - temp) * prcp",
                           "rainfall ~ step_func(temp - Tmin) * prcp"], symbols=s),
                HydroFlux("melt ~ step_func(temp - Tmax) * step_func(snowpack) * min(snowpack, Df * (temp - Tmax))",
                          symbols=s),
                HydroFlux(PET_HAMON, symbols=s),
                HydroFlux("evap ~ step_func(soilwater) * pet * min(1.0, soilwater / Smax)", symbols=s),
                HydroFlux("baseflow ~ step_func(soilwater) * Qmax * exp(-f * (max(0.0, Smax - soilwater)))",
                          symbols=s),
                HydroFlux("surfaceflow ~ max(0.0, soilwater - Smax)", symbols=s),
                HydroFlux("flow ~ baseflow + surfaceflow", symbols=s)],
        dfluxes=[StateFlux("snowpack ~ snowfall - melt", symbols=s),
                 StateFlux("soilwater ~ (rainfall + melt) - (evap + flow)", symbols=s)], htypes=htypes)
    return HydroModel(name="exphydro_v2", components=(b,), input_names=["temp", "lday", "prcp"])


def gr4j(htypes=None) -> HydroModel:
    """GR4J with two unit hydrographs.  Inputs `prcp, ep`; params `x1..x4`; 15 rows."""
    s = Symbols("prcp ep soilwater pn en ps es perc pr slowflow fastflow t slowflow_routed fastflow_routed "
                "routingstore exch routedflow flow", "x1 x2 x3 x4")
    prod = HydroBucket(
        name="gr4j_prod",
        fluxes=[HydroFlux("pn ~ prcp - min(prcp, ep)", symbols=s),
                HydroFlux("en ~ ep - min(prcp, ep)", symbols=s),
                HydroFlux("ps ~ max(0.0, pn * (1 - (soilwater / x1)^2))", symbols=s),
                HydroFlux("es ~ en * (2 * soilwater / x1 - (soilwater / x1)^2)", symbols=s),
                HydroFlux("perc ~ ((x1)^(-4)) / 4 * ((4 / 9)^(4)) * (soilwater^5)", symbols=s),
                HydroFlux("pr ~ pn - ps + perc", symbols=s),
                HydroFlux("slowflow ~ 0.9 * pr", symbols=s),
                HydroFlux("fastflow ~ 0.1 * pr", symbols=s)],
        dfluxes=[StateFlux("soilwater ~ ps - es - perc", symbols=s)], htypes=htypes)
    uh_slow = UnitHydrograph(s.slowflow, s.slowflow_routed, [("x4", "(t / x4)^2.5")],
                             symbols=s, htypes=htypes, name="uh_slow")
    uh_fast = UnitHydrograph(s.fastflow, s.fastflow_routed,
                             [("2x4", "(1 - 0.5 * (2 - t / x4)^2.5)"), ("x4", "(0.5 * (t / x4)^2.5)")],
                             symbols=s, htypes=htypes, name="uh_fast")
    rout = HydroBucket(
        name="gr4j_rout",
        fluxes=[HydroFlux("exch ~ x2 * abs(routingstore / x3)^3.5", symbols=s),
                HydroFlux("routedflow ~ x3^(-4) / 4 * (routingstore + slowflow_routed + exch)^5", symbols=s),
                HydroFlux("flow ~ routedflow + max(fastflow_routed + exch, 0.0)", symbols=s)],
        dfluxes=[StateFlux("routingstore ~ slowflow_routed + exch - routedflow", symbols=s)], htypes=htypes)
    return HydroModel(name="gr4j", components=(prod, uh_slow, uh_fast, rout), input_names=["prcp", "ep"])


def hbv(htypes=None) -> HydroModel:
    """HBV (snow / soil / response zones), 5 states, last output `q`.  Inputs
    `prcp, temp, pet`; params `TT CFMAX CFR CWH LP FC BETA PPERC UZL k0 k1 k2`."""
    s = Symbols("soilwater snowpack meltwater suz slz prcp pet temp rainfall snowfall melt refreeze infil "
                "excess recharge evap q0 q1 q2 q perc", "TT CFMAX CFR CWH LP FC BETA PPERC UZL k0 k1 k2")
    split = HydroFlux(["snowfall ~ step_func(TT - temp) * prcp", "rainfall ~ step_func(temp - TT) * prcp"],
                      symbols=s, htypes=htypes)
    snow = HydroBucket(
        name="hbv_snow",
        fluxes=[HydroFlux("melt ~ min(snowpack, max(0.0, temp - TT) * CFMAX)", symbols=s),
                HydroFlux("refreeze ~ min(max((TT - temp), 0.0) * CFR * CFMAX, meltwater)", symbols=s),
                HydroFlux("infil ~ max(0.0, meltwater - snowpack * CWH)", symbols=s)],
        dfluxes=[StateFlux("snowpack ~ snowfall + refreeze - melt", symbols=s),
                 StateFlux("meltwater ~ melt - refreeze - infil", symbols=s)], htypes=htypes)
    soil = HydroBucket(
        name="hbv_soil",
        fluxes=[HydroFlux("recharge ~ (rainfall + infil) * clamp((max(soilwater / FC, 0.0))^(BETA * 5 + 1), 0, 1)",
                          symbols=s),
                HydroFlux("excess ~ max(soilwater - FC, 0.0)", symbols=s),
                HydroFlux("evap ~ clamp(soilwater / (LP * FC), 0, 1) * pet", symbols=s)],
        dfluxes=[StateFlux("soilwater ~ (rainfall + infil) - (recharge + excess + evap)", symbols=s)],
        htypes=htypes)
    zone = HydroBucket(
        name="hbv_zone",
        fluxes=[HydroFlux("perc ~ suz * PPERC", symbols=s),
                HydroFlux("q0 ~ max(0.0, suz - UZL) * k0", symbols=s),
                HydroFlux("q1 ~ suz * k1", symbols=s),
                HydroFlux("q2 ~ slz * k2", symbols=s),
                HydroFlux("q ~ q0 + q1 + q2", symbols=s)],
        dfluxes=[StateFlux("suz ~ recharge + excess - (perc + q0 + q1)", symbols=s),
                 StateFlux("slz ~ perc - q2", symbols=s)], htypes=htypes)
    return HydroModel(name="hbv", components=(split, snow, soil, zone), input_names=["prcp", "temp", "pet"])


def m50_chains():
    et_nn = Chain((Dense(3, 16, "tanh"), Dense(16, 16, "leakyrelu"), Dense(16, 1, "leakyrelu")), name="etnn")
    q_nn = Chain((Dense(2, 16, "tanh"), Dense(16, 16, "leakyrelu"), Dense(16, 1, "leakyrelu")), name="qnn")
    return et_nn, q_nn


def m50(htypes=None) -> HydroModel:
    """M50: ExpHydro snow bucket + a soil bucket whose ET and Q fluxes are two Lux MLPs
    (3-16-16-1 and 2-16-16-1).  Inputs `prcp, temp, lday`."""
    s = Symbols("prcp temp lday pet rainfall snowfall snowpack soilwater melt log_evap_div_lday log_flow "
                "norm_snw norm_slw norm_temp norm_prcp",
                "Tmin Tmax Df snowpack_std snowpack_mean soilwater_std soilwater_mean prcp_std prcp_mean "
                "temp_std temp_mean")
    snow = HydroBucket(
        name="m50_snow",
        fluxes=[HydroFlux(PET_HAMON, symbols=s),
                HydroFlux("snowfall ~ step_func(Tmin - temp) * prcp", symbols=s),
                HydroFlux("rainfall ~ step_func(temp - Tmin) * prcp", symbols=s),
                HydroFlux("melt ~ step_func(temp - Tmax) * step_func(snowpack) * min(snowpack, Df * (temp - Tmax))",
                          symbols=s)],
        dfluxes=[StateFlux("snowpack ~ snowfall - melt", symbols=s)], htypes=htypes)
    et_nn, q_nn = m50_chains()
    soil = HydroBucket(
        name="m50_soil",
        fluxes=[HydroFlux("norm_snw ~ (snowpack - snowpack_mean) / snowpack_std", symbols=s),
                HydroFlux("norm_slw ~ (soilwater - soilwater_mean) / soilwater_std", symbols=s),
                HydroFlux("norm_prcp ~ (prcp - prcp_mean) / prcp_std", symbols=s),
                HydroFlux("norm_temp ~ (temp - temp_mean) / temp_std", symbols=s),
                NeuralFlux([s.norm_snw, s.norm_slw, s.norm_temp], s.log_evap_div_lday, et_nn),
                NeuralFlux([s.norm_slw, s.norm_prcp], s.log_flow, q_nn)],
        dfluxes=[StateFlux("soilwater ~ rainfall + melt - step_func(soilwater) * lday * log_evap_div_lday "
                           "- step_func(soilwater) * exp(log_flow)", symbols=s)], htypes=htypes)
    return HydroModel(name="m50", components=(snow, soil), input_names=["prcp", "temp", "lday"])


def m50_three_components(htypes=None) -> HydroModel:
    """The M50 of `/root/reference/test/models/m50.jl`: snow bucket, soil bucket whose two neural fluxes give
    `log(evap / lday)` and `log(flow)` (both exponentiated in the soil balance, `:72`), and a third, stateless component
    `flow ~ exp(log_flow)` (`:77`).  Inputs `prcp, temp, lday`; chains `epnn` (3-16-16-1) and `qnn` (2-16-16-1)."""
    s = Symbols("prcp temp lday pet rainfall snowfall snowpack soilwater melt log_evap_div_lday log_flow flow "
                "norm_snw norm_slw norm_temp norm_prcp",
                "Tmin Tmax Df snowpack_std snowpack_mean soilwater_std soilwater_mean prcp_std prcp_mean "
                "temp_std temp_mean")
    snow = HydroBucket(
        name="m50_snow",
        fluxes=[HydroFlux(PET_HAMON, symbols=s),
                HydroFlux(["snowfall ~ step_func(Tmin - temp) * prcp", "rainfall ~ step_func(temp - Tmin) * prcp"], symbols=s),
                HydroFlux("melt ~ step_func(temp - Tmax) * min(snowpack, Df * (temp - Tmax))", symbols=s)],
        dfluxes=[StateFlux("snowpack ~ snowfall - melt", symbols=s)], htypes=htypes)
    ep_nn = Chain((Dense(3, 16, "tanh"), Dense(16, 16, "leakyrelu"), Dense(16, 1, "leakyrelu")), name="epnn")
    q_nn = Chain((Dense(2, 16, "tanh"), Dense(16, 16, "leakyrelu"), Dense(16, 1, "leakyrelu")), name="qnn")
    soil = HydroBucket(
        name="m50_soil",
        fluxes=[HydroFlux("norm_snw ~ (snowpack - snowpack_mean) / snowpack_std", symbols=s),
                HydroFlux("norm_slw ~ (soilwater - soilwater_mean) / soilwater_std", symbols=s),
                HydroFlux("norm_prcp ~ (prcp - prcp_mean) / prcp_std", symbols=s),
                HydroFlux("norm_temp ~ (temp - temp_mean) / temp_std", symbols=s),
                NeuralFlux([s.norm_snw, s.norm_slw, s.norm_temp], s.log_evap_div_lday, ep_nn),
                NeuralFlux([s.norm_slw, s.norm_prcp], s.log_flow, q_nn)],
        dfluxes=[StateFlux("soilwater ~ rainfall + melt - step_func(soilwater) * lday * exp(log_evap_div_lday) "
                           "- step_func(soilwater) * exp(log_flow)", symbols=s)], htypes=htypes)
    convert = HydroFlux("flow ~ exp(log_flow)", symbols=s, htypes=htypes)
    return HydroModel(name="m50", components=(snow, soil, convert), input_names=["prcp", "temp", "lday"])


def exphydro_grid(flwdir, positions, htypes: Optional[Sequence[int]] = None) -> HydroModel:
    """Distributed ExpHydro + `q ~ flow * area_coef` + D8 grid route
    (`q_routed ~ s_river / (1 + lag) + q`, `s_river ~ q - q_routed`)."""
    n = len(positions)
    ht = np.arange(1, n + 1) if htypes is None else np.asarray(htypes)
    _, b1, b2 = _exphydro_buckets(ht, docs_variant=False)
    s = Symbols("flow q q_routed s_river", "area_coef lag")
    convert = HydroFlux("q ~ flow * area_coef", symbols=s, htypes=ht)
    route = HydroRoute(rfluxes=[HydroFlux("q_routed ~ s_river / (1 + lag) + q", symbols=s)],
                       dfluxes=[StateFlux("s_river ~ q - q_routed", symbols=s)],
                       aggr_func=build_aggr_func(flwdir, positions), htypes=ht, name="grid_route")
    return HydroModel(name="exphydro_grid", components=(b1, b2, convert, route),
                      input_names=["temp", "lday", "prcp"])
