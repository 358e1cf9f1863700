"""helper of tests/test_gpu_mpi_dropin.py and tests/test_minimpi_reference.py: writes the parameter file (the reference's own
nbody/parameter.txt values unless overridden) and a format-1 initial-conditions file for a run of the reference's REAL
main() on several tasks (oracle/_ref/sidm_ref_mpi = all CPU, oracle/_ref/sidm_b200_mpi = GPU drop-in), and starts it."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "sidm-nbody_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

PARAMS = dict(
    InitCondFile="ic.dat", OutputDir="./", EnergyFile="energy_out", InfoFile="info_out", TimingsFile="timings_out", CpuFile="cpu_out",
    RestartFile="rst_out", SnapshotFileBase="snp", CrossSection=0.0, RandomSeed1=55, RandomSeed2=497527, ProbabilityTol=0.2,
    ReflectionBoundary=1114.35, TimeLimitCPU=86400.0, ResubmitOn=0, ResubmitCommand="xyz", ICFormat=1, ComovingIntegrationOn=0,
    NumFilesPerSnapshot=1, NumFilesWrittenInParallel=1, CoolingOn=0, TimeBegin=0.0, TimeMax=0.01, Omega0=1.0, OmegaLambda=0.0,
    OmegaBaryon=0.0, HubbleParam=0.7, BoxSize=0.0, PeriodicBoundariesOn=0, OutputListFilename="lst_in", OutputListOn=0,
    TimeBetSnapshot=1000.0, TimeOfFirstSnapshot=1000.0, CpuTimeBetRestartFile=72000.0, TimeBetStatistics=1000.0, TypeOfTimestepCriterion=1,
    ErrTolIntAccuracy=0.005, ErrTolDynamicalAccuracy=0.004, ErrTolVelScale=0.66, MaxSizeTimestep=100.0, MinSizeTimestep=0.0,
    ErrTolTheta=0.5, TypeOfOpeningCriterion=1, ErrTolForceAcc=0.005, MaxNodeMove=0.02, TreeUpdateFrequency=0.1, DesNumNgb=30,
    MaxNumNgbDeviation=2, ArtBulkViscConst=0.75, InitGasTemp=1000.0, MinGasTemp=1000.0, CourantFac=0.15, PartAllocFactor=2.0,
    TreeAllocFactor=0.8, BufferSize=30, DomainUpdateFrequency=0.5, UnitLength_in_cm=3.085678e21, UnitMass_in_g=1.989e43,
    UnitVelocity_in_cm_per_s=1e5, GravityConstantInternal=0, MinGasHsmlFractional=1.0, SofteningGas=600.0, SofteningHalo=0.3,
    SofteningDisk=0.0, SofteningBulge=0.0, SofteningStars=0.0, SofteningGasMaxPhys=200.0, SofteningHaloMaxPhys=0.3,
    SofteningDiskMaxPhys=0.0, SofteningBulgeMaxPhys=0.0, SofteningStarsMaxPhys=0.0)


def write_case(workdir, n, seed=21, **over):
    import oracle
    from sidm_b200 import ic
    os.makedirs(workdir, exist_ok=True)
    pos, vel, mass, ids = ic.hernquist(n, seed=seed)
    par = dict(PARAMS)
    par.update(over)
    nf = int(par["NumFilesPerSnapshot"])
    if nf <= 1:
        with open(os.path.join(workdir, "ic.dat"), "wb") as f:
            f.write(oracle.snapshot_bytes(pos, vel, ids, mass))
    else:                                               # read_ic.c:62-75: the initial conditions are split like the snapshots
        cuts = np.linspace(0, n, nf + 1).astype(int)
        for k in range(nf):
            a, b = cuts[k], cuts[k + 1]
            with open(os.path.join(workdir, f"ic.dat.{k}"), "wb") as f:
                f.write(oracle.snapshot_bytes(pos[a:b], vel[a:b], ids[a:b], mass[a:b], npart_total=[0, n, 0, 0, 0, 0], num_files=nf))
    with open(os.path.join(workdir, "param.txt"), "w") as f:
        for k, v in par.items():
            f.write(f"{k:28s} {v}\n")
        f.write("\n")                                   # the reference's parser needs a final blank line
    return pos, vel, mass, ids


def run_case(workdir, exe, ntask, timeout=900, env=None):
    e = dict(os.environ, MINIMPI_NP=str(ntask))
    if env:
        e.update(env)
    r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", exe), "param.txt"], cwd=workdir, capture_output=True, text=True, timeout=timeout, env=e)
    return r


def last_snapshot(workdir, base="snp"):
    import oracle
    files = sorted(f for f in os.listdir(workdir) if f.startswith(base + "_"))
    assert files, os.listdir(workdir)
    s = oracle.read_snapshot(os.path.join(workdir, files[-1]))
    o = np.argsort(s["ids"])
    return dict(time=s["time"], ids=s["ids"][o], pos=s["pos"][o], vel=s["vel"][o], files=files)
