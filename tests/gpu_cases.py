"""Inputs of the `-m gpu` parity tests, in ONE place, so that a CPU-collected test (tests/test_gpu_inputs_build.py) can build
every one of them without a GPU: a GPU test whose inputs cannot even be generated must fail here, in the container, not on
the driver's B200 run (round 1 shipped a test that asked for 3 neurons on a 20 x 16 field with an 8-px margin).

GPU tests take their synthetic videos from `synthetic(name)` / `traces(name)` only; tests/test_gpu_inputs_build.py also checks
that no GPU test module calls the generators directly."""
from oracle import gen, oasis as O

# name -> make_synthetic keyword arguments (oracle/gen.py)
SYNTHETIC = {
    "chain_64": dict(d1=64, d2=64, T=1000, K=6, seed=7, nblob=4),
    "chain_96x80_patches": dict(d1=96, d2=80, T=600, K=10, seed=7, nblob=4),
    "hals_nodeconv": dict(d1=64, d2=64, T=800, K=5, seed=11, nblob=4),
    "bg_identity": dict(d1=128, d2=128, T=1200, K=12, seed=3),
    "sn_lars": dict(d1=64, d2=48, T=700, K=5, seed=21, nblob=4),
    "fast_temporal": dict(d1=64, d2=48, T=600, K=5, seed=33, nblob=4),
    "ring18": dict(d1=56, d2=48, T=500, K=4, seed=5, nblob=4),
    "no_neurons": dict(d1=40, d2=36, T=300, K=3, seed=5, nblob=2),
    "noise": dict(d1=40, d2=36, T=400, K=3, seed=8, nblob=2),
    "ssub_60x52": dict(d1=60, d2=52, T=500, K=5, seed=41, nblob=3),
    "ssub_75x64_patches": dict(d1=75, d2=64, T=400, K=7, seed=41, nblob=3),
    "ssub3_48x45": dict(d1=48, d2=45, T=300, K=4, seed=41, nblob=3),
    "svd_48x40": dict(d1=48, d2=40, T=500, K=5, seed=31, nblob=3, bg_amp=60.0),
    "svd_64x60_patches": dict(d1=64, d2=60, T=400, K=8, seed=31, nblob=3, bg_amp=60.0),
    "nmf_sub": dict(d1=40, d2=36, T=400, K=4, seed=8, nblob=2, bg_amp=50.0),
    "nmf_fit": dict(d1=48, d2=40, T=500, K=5, seed=17, nblob=3, bg_amp=60.0),
    "hals_uv": dict(d1=48, d2=48, T=1500, K=8, seed=5, nblob=0, bg_amp=0.0),
    "outlier": dict(d1=56, d2=48, T=500, K=4, seed=5, nblob=4),
    "kf2_long": dict(d1=48, d2=40, T=4200, K=4, seed=9, nblob=3),
    "multi_gpu": dict(d1=96, d2=80, T=600, K=10, seed=7, nblob=4),
    "post_dev": dict(d1=64, d2=64, T=600, K=6, seed=12, nblob=3),
}
NOISE_CASE = (40, 36, 400, 3, 6)      # d1, d2, T, K, ring radius of "noise"

# name -> (gen_data arguments g, noise, T, framerate, firerate, b, N, seed) (oracle/oasis.py gen_data)
TRACES = {
    "ar1_n01": (0.95, 0.1, 3000, 30, 0.5, 0, 6, 13),
    "ar1_n03": (0.95, 0.3, 3000, 30, 0.5, 0, 6, 13),
    "ar1_baseline": (0.95, 0.3, 3000, 30, 0.5, 0.7, 5, 17),
    "ar1_getsn_1000": (0.95, 0.3, 1000, 30, 0.5, 0, 3, 13),
    "ar1_getsn_3000": (0.95, 0.3, 3000, 30, 0.5, 0, 3, 13),
    "ar1_getsn_10000": (0.95, 0.3, 10000, 30, 0.5, 0, 3, 13),
    "ar1_getsn_20011": (0.95, 0.3, 20011, 30, 0.5, 0, 3, 13),
    "ar1_long": (0.95, 0.2, 100000, 30, 0.5, 0, 8, 13),
    "ar2_4": ([1.7, -0.712], 1.0, 3000, 30, 0.5, 0, 4, 3),
    "ar2_3": ([1.7, -0.712], 1.0, 3000, 30, 0.5, 0, 3, 3),
    "ar2_c5": ([1.7, -0.712], 1.0, 100000, 30, 0.5, 0, 8, 3),
}


def synthetic(name):
    D = gen.make_synthetic(**SYNTHETIC[name])
    assert D["A0"].shape[1] == SYNTHETIC[name]["K"], "case %s: only %d of %d neurons fit the field" % (
        name, D["A0"].shape[1], SYNTHETIC[name]["K"])
    return D


def traces(name):
    g, noise, T, fr, rate, b, N, seed = TRACES[name]
    return O.gen_data(g, noise, T, fr, rate, b, N, seed)
