"""The five BASELINE.json workloads as reference command lines.

Assembled exactly as the reference's scripts/do-fundamentals.sh:396-419 does from the
variables each do-script sets (do-halfspace.sh, do-halfspace-nearsrc50.sh,
do-crustpinch.sh, do-lopnor.sh, do-spherical.sh), with only --num-phonons,
--toa-degree and --output-dir left open.  Used by the golden-fixture generator,
by the tests, and by bench.py / the integration binary to build the model of a workload.
"""

_HALFSPACE_ARGS = "0.8,0.01,1.0,0.5,1000,0.8,0.01,1.0,0.5,1000,6.40,3.63,2.83,-60,6.40,3.63,2.83,-400"
_CP_SCAT = "0.8,0.01,0.20,0.2,200,0.8,0.01,0.20,0.3,1500,0.8,0.01,0.20,0.3,1500,0.8,0.01,0.20,0.4,1500,0.8,0.01,0.20,0.5,900"

CONFIGS = {
    # do-halfspace.sh
    "halfspace": dict(
        source="SDR,0,90,0,0.0", source_loc="0,0,-5", frequency="2.0", ttl="200", binsize="0.50",
        grid=40, rng="900", flatten=False, model_args=_HALFSPACE_ARGS,
        seis=["0,0,0,183.85,183.85,0,2.737,0.105,10.0,48", "0,0,0,260,0,0,2.737,0.105,10.0,48",
              "0,0,0,240.21,-99.5,0,2.737,0.105,10.0,48"]),
    # do-halfspace-nearsrc50.sh
    "halfspace_nearsrc50": dict(
        source="SDR,0,90,0,0.0", source_loc="0,0,-5", frequency="2.0", ttl="125", binsize="0.10",
        grid=40, rng="700", flatten=False, model_args=_HALFSPACE_ARGS,
        seis=["0,0,0,35.356,35.356,0,0.5263,0.0211,2.0,48", "0,0,0,50,0,0,0.5263,0.0211,2.0,48",
              "0,0,0,46.194,-19.135,0,0.5263,0.0211,2.0,48"]),
    # do-crustpinch.sh
    "crustpinch": dict(
        source="SDR,22.5,90,0", source_loc="0,0,-10", frequency="2.0", ttl="600", binsize="2.00",
        grid=5, rng=None, flatten=False, model_args=_CP_SCAT + ",2.0,30.0,5.0,.3666667,.4736842,1,1",
        seis=["0,67.5,0,950,67.5,0,1.0,2.0,40.0,160", "0,112.5,0,950,112.5,0,1.0,2.0,40.0,160",
              "0,90,0,950,90,0,1.0,2.0,40.0,160"]),
    # do-lopnor.sh (15-argument baseline model)
    "lopnor": dict(
        source="SDR,125,40,90,0.0", source_loc="425.54,-169.53,-31.02", frequency="2.0", ttl="600", binsize="2.00",
        grid=1, rng="1200", flatten=True, model_args="0.8,0.01,0.5,0.2,50,0.8,0.01,0.5,0.3,1000,0.8,0.01,0.7,0.5,300",
        seis=["425.54,-169.53,0.98,-390.04,-167.18,1.457,1.0,2.0,40.0,160",
              "425.54,-169.53,0.98,-102.27,430.84,0.60,1.0,2.0,40.0,160"]),
    # do-spherical.sh
    "spherical": dict(
        source="SDR,22.5,90,0", source_loc="0,0,-10", frequency="2.0", ttl="8000", binsize="20.0",
        grid=16, rng=None, flatten=False, model_args=_CP_SCAT,
        seis=["0,67.5,0,12000,67.5,0,20.0,20.0,400.0,160", "0,112.5,0,12000,112.5,0,20.0,20.0,400.0,160",
              "0,90,0,12000,90,0,20.0,20.0,400.0,160"]),
}


def cmdline(config, n_phonons, toa_degree, outdir, extra=()):
    c = CONFIGS[config]
    a = ["--reports=INV", f"--output-dir={outdir}", "--report-file=reports.dat",
         f"--num-phonons={n_phonons}", f"--toa-degree={toa_degree}", f"--source={c['source']}",
         f"--source-loc={c['source_loc']}", f"--frequency={c['frequency']}", f"--timetolive={c['ttl']}",
         f"--binsize={c['binsize']}", f"--grid-compiled={c['grid']}"]
    if c["rng"]:
        a.append(f"--range={c['rng']}")
    if c["flatten"]:
        a.append("--flatten")
    a.append(f"--model-args={c['model_args']}")
    a += [f"--seis-p2p={s}" for s in c["seis"]]
    a += list(extra)
    return a
