"""User-model SDK: an out-of-tree .cu compiled against the header-only kernel templates registers itself at dlopen time
(examples/plugin_model) and runs through the same engine as the built-in models."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import sabc_b200 as sb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN_DIR = os.path.join(ROOT, "examples", "plugin_model")


@pytest.fixture(scope="module")
def plugin():
    lib = os.path.join(PLUGIN_DIR, "libsabc_ar1.so")
    src = os.path.join(PLUGIN_DIR, "ar1_model.cu")
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(src), os.path.getmtime(sb._lib.LIB_PATH)):
        subprocess.run([os.path.join(PLUGIN_DIR, "build.sh")], check=True, capture_output=True)
    sb._lib.lib()                                        # the engine library first (RTLD_GLOBAL), then the plug-in
    return C.CDLL(lib, mode=C.RTLD_GLOBAL)


def test_plugin_registers_itself(plugin):
    assert "ar1" in sb.models.registered()
    d, s = C.c_int32(), C.c_int32()
    assert sb._lib.lib().sabc_model_info(b"ar1", C.byref(d), C.byref(s)) == 0 and (d.value, s.value) == (2, 2)


def ar1_reference(theta, par, seed, particle, sweep):
    """the plug-in's arithmetic restated with the oracle's Philox / normal primitives"""
    import oracle_binding as ob
    T = int(par[0]); x = s1 = s2 = sx = 0.0
    zs = np.zeros(2 * ((T + 1) // 2))
    ob.lib().orc_normal_stream(seed, particle, sweep, zs.size // 2, ob.p(zs))     # the model stream's ziggurat normals, in order
    for t in range(0, T, 2):
        for h, z in enumerate((zs[t], zs[t + 1])):
            if t + h < T:
                xn = theta[0] * x + theta[1] * z
                s1 = s1 + xn * x; s2 = s2 + xn * xn; sx = sx + xn; x = xn
    m = sx / T
    return abs(s1 / T - par[1]), abs((s2 / T - m * m) - par[2])


@pytest.mark.gpu
def test_plugin_runs_on_the_engine(gpu, plugin):
    par = np.array([40.0, 0.35, 0.9])
    model = sb.DeviceModel("ar1", 2, 2, par)
    rng = np.random.default_rng(0)
    th = np.column_stack([rng.uniform(-0.9, 0.9, 200), rng.uniform(0.1, 2, 200)])
    rho = model.simulate(th, seed=5, particle_base=10, sweep=3)
    for i in range(200):
        want = ar1_reference(th[i], par, 5, 10 + i, 3)
        assert rho[i, 0] == want[0] and rho[i, 1] == want[1]
    prior = sb.product_distribution([sb.Uniform(-0.95, 0.95), sb.Uniform(0.05, 3.0)])
    for algorithm in ("single_eps", "multi_eps"):
        res = sb.sabc(model, prior, n_particles=2000, n_simulation=60_000, algorithm=algorithm)
        assert res.state.n_population_updates == 29 and np.all(res.state.eps < 0.5) and np.all(np.isfinite(res.population))
        assert abs(np.median(res.population[:, 0]) - 0.4) < 0.35       # obs correspond to phi ~ 0.4, sigma ~ 0.9


def test_register_rejects_a_foreign_launch_table(plugin):
    """the launch table carries its own size and version: a plug-in compiled against other kernel headers is refused, not called"""
    bogus = (C.c_uint32 * 64)()
    bogus[0], bogus[1] = 8, 1
    assert sb._lib.lib().sabc_register_model(bogus) == -20
    assert b"launch table" in sb._lib.lib().sabc_last_error()


@pytest.mark.gpu
def test_engine_survives_a_plugin_loaded_later(gpu):
    """an engine keeps a pointer to its model's registry entry; loading a plug-in afterwards (dlopen -> sabc_register_model) must not
    move that entry.  Separate process: the plug-in must not be loaded yet when the engine is created."""
    code = f"""
import ctypes as C, sys
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
import numpy as np, sabc_b200 as sb
kw = dict(n_particles=2000, algorithm="single_eps", proposal=sb.DifferentialEvolution(n_para=1), resample=4000, v=1.0, delta=0.1)
a = sb.Engine(sb.models.gauss_mean(1.0), sb.Normal(0, 1), **kw); a.init(); a.update(5 * 2000)
n0 = sb._lib.lib().sabc_model_count()
C.CDLL({os.path.join(PLUGIN_DIR, 'libsabc_ar1.so')!r}, mode=C.RTLD_GLOBAL)
assert sb._lib.lib().sabc_model_count() == n0 + 1
a.update(5 * 2000)
b = sb.Engine(sb.models.gauss_mean(1.0), sb.Normal(0, 1), **kw); b.init(); b.update(10 * 2000)
assert all(np.array_equal(x, y) for x, y in zip(a.get_population(), b.get_population()))
print("ok")
"""
    r = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
