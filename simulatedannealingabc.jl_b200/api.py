"""Host-side mirror of the reference's public surface for the hot path:

    sabc(f_dist, prior, args...; n_particles, n_simulation, algorithm, proposal, resample, v, δ,
         checkpoint_history, show_progressbar, show_checkpoint, kwargs...) -> SABCresult      (src/SimulatedAnnealingABC.jl:451-492)
    update_population!(res, f_dist, prior, args...; n_simulation, v, δ, proposal, resample, ...)  (:251-402)
    SABCresult(population, u, ρ, state), SABCstate(ϵ, algorithm, histories, counters)            (:28-60)

`f_dist` is a DeviceModel (models.py) in place of a closure; arbitrary closures stay on the reference's CPU path and
are rejected here.  All computation happens in libsabc_b200.so; this file only validates arguments like the
reference does and moves arrays across the C ABI."""
from __future__ import annotations

import ctypes as C
import math
import warnings

import numpy as np

from . import _lib
from .distributions import Distribution
from .models import DeviceModel
from .proposals import DifferentialEvolution, Proposal, RandomWalk, StretchMove  # noqa: F401

ALGORITHMS = {"single_eps": 0, "multi_eps": 1}
# north_star vocabulary: type=:single|:multi|:hybrid  (SURVEY.md §0).  The reference (v0.4.0) has algorithm = :single_eps | :multi_eps
# only; "hybrid" is its :single_eps applied to several statistics (one eps, per-statistic ECDFs, sum_j du_j / eps,
# src/SimulatedAnnealingABC.jl:318-319, table :439-446) -- state.algorithm therefore reports single_eps for it.
TYPES = {"single": "single_eps", "multi": "multi_eps", "hybrid": "single_eps"}


class Engine:
    """Owns one sabc_engine handle (device memory, stream, communicator)."""

    def __init__(self, model: DeviceModel, prior: Distribution, *, n_particles: int, algorithm: str, proposal: Proposal,
                 resample: int, v: float, delta: float, seed: int = 0x5ABC, device: int = -1, rank: int = 0,
                 world_size: int = 1, nccl_unique_id: bytes | None = None, flags: int = 0, ecdf_max_knots: int = 0,
                 n_gpus: int = 0, gpu_ids=None):
        comps = prior.components()
        if len(comps) != model.n_para:
            raise _lib.SABCError(-20, f"prior has {len(comps)} components but model '{model.name}' has {model.n_para} parameters")
        self.model, self.prior = model, prior
        self.N, self.d, self.s = int(n_particles), model.n_para, model.n_stats
        self.n_eps = self.s if algorithm == "multi_eps" else 1
        self._par = np.ascontiguousarray(model.par, dtype=np.float64)
        self._kind = np.array([c.kind for c in comps], dtype=np.int32)
        self._ppar = np.array([p for c in comps for p in c.params()], dtype=np.float64)
        self._uid = C.create_string_buffer(nccl_unique_id, 128) if nccl_unique_id is not None else None
        cfg = _lib.Config()
        cfg.n_particles = self.N
        cfg.n_para, cfg.n_stats = self.d, self.s
        cfg.algorithm = ALGORITHMS[algorithm]
        cfg.proposal = proposal.kind
        cfg.prop_par[0], cfg.prop_par[1] = proposal.params()
        cfg.v, cfg.delta, cfg.resample, cfg.seed = float(v), float(delta), int(resample), int(seed)
        cfg.model_name = model.name.encode()
        cfg.model_par = self._par.ctypes.data_as(_lib.c_double_p)
        cfg.n_model_par = self._par.size
        cfg.device = device
        cfg.prior_kind = self._kind.ctypes.data_as(_lib.c_int32_p)
        cfg.prior_par = self._ppar.ctypes.data_as(_lib.c_double_p)
        cfg.rank, cfg.world_size = rank, world_size
        cfg.nccl_unique_id = C.cast(self._uid, C.c_void_p) if self._uid is not None else None
        cfg.flags = flags
        cfg.ecdf_max_knots = int(ecdf_max_knots)
        # one process, several GPUs: the handle shards the particles itself (host arrays stay the global N x d / N x s matrices)
        self._gpu_ids = np.ascontiguousarray(gpu_ids, dtype=np.int32) if gpu_ids is not None else None
        cfg.n_gpus = int(n_gpus) if self._gpu_ids is None else int(self._gpu_ids.size)
        cfg.gpu_ids = self._gpu_ids.ctypes.data_as(_lib.c_int32_p) if self._gpu_ids is not None else None
        self._h = C.c_void_p()
        _lib.check(_lib.lib().sabc_create(C.byref(self._h), C.byref(cfg)))
        nl, off = C.c_int64(), C.c_int64()
        _lib.check(_lib.lib().sabc_local_particles(self._h, C.byref(nl), C.byref(off)))
        self.n_local, self.offset = nl.value, off.value

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().sabc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- the path ---
    def init(self):
        _lib.check(_lib.lib().sabc_init(self._h))

    def update(self, n_simulation: int, checkpoint_history: int = 1):
        _lib.check(_lib.lib().sabc_update(self._h, int(n_simulation), int(checkpoint_history)))

    def update_host(self, theta, u, rho, eps, counters, n_simulation: int, checkpoint_history: int = 1):
        _lib.check(_lib.lib().sabc_update_host(self._h, _lib.ptr(theta), _lib.ptr(u), _lib.ptr(rho), _lib.ptr(eps),
                                               _lib.ptr(counters), int(n_simulation), int(checkpoint_history)))

    def set_tuning(self, v, delta, resample, proposal: Proposal):
        pp = (C.c_double * 2)(*proposal.params())
        _lib.check(_lib.lib().sabc_set_tuning(self._h, float(v), float(delta), int(resample), proposal.kind, pp))

    # --- state ---
    def get_population(self, theta=True, u=True, rho=True):
        n = self.n_local
        th = np.empty((n, self.d), order="F") if theta else None
        uu = np.empty((n, self.s), order="F") if u else None
        rr = np.empty((n, self.s), order="F") if rho else None
        _lib.check(_lib.lib().sabc_get_population(self._h, _lib.ptr(th), _lib.ptr(uu), _lib.ptr(rr)))
        return th, uu, rr

    def set_population(self, theta, u, rho, eps, counters):
        th = _lib.f64(np.asarray(theta, dtype=np.float64).reshape(self.n_local, self.d, order="F"))
        uu = _lib.f64(np.asarray(u, dtype=np.float64).reshape(self.n_local, self.s, order="F"))
        rr = _lib.f64(np.asarray(rho, dtype=np.float64).reshape(self.n_local, self.s, order="F"))
        ee = np.ascontiguousarray(eps, dtype=np.float64)
        cc = np.ascontiguousarray(counters, dtype=np.int64)
        _lib.check(_lib.lib().sabc_set_population(self._h, _lib.ptr(th), _lib.ptr(uu), _lib.ptr(rr), _lib.ptr(ee), _lib.ptr(cc)))

    def get_state(self):
        eps = np.zeros(self.n_eps)
        cnt = np.zeros(4, dtype=np.int64)
        _lib.check(_lib.lib().sabc_get_state(self._h, _lib.ptr(eps), _lib.ptr(cnt)))
        return eps, cnt

    def get_history(self):
        n = C.c_int64()
        _lib.check(_lib.lib().sabc_history_len(self._h, C.byref(n)))
        e = np.zeros((n.value, self.n_eps)); u = np.zeros((n.value, self.s)); r = np.zeros((n.value, self.s))
        _lib.check(_lib.lib().sabc_get_history(self._h, _lib.ptr(e), _lib.ptr(u), _lib.ptr(r)))
        return e, u, r

    def get_ecdf(self, stat: int) -> np.ndarray:
        L = C.c_int64()
        _lib.check(_lib.lib().sabc_get_ecdf(self._h, stat, None, C.byref(L)))
        k = np.zeros(L.value)
        _lib.check(_lib.lib().sabc_get_ecdf(self._h, stat, _lib.ptr(k), C.byref(L)))
        return k

    def set_ecdf(self, stat: int, knots):
        k = np.ascontiguousarray(knots, dtype=np.float64)
        _lib.check(_lib.lib().sabc_set_ecdf(self._h, stat, _lib.ptr(k), k.size))

    def timing(self) -> dict:
        t = _lib.Timing()
        _lib.check(_lib.lib().sabc_get_timing(self._h, C.byref(t)))
        return {f: getattr(t, f) for f, _ in t._fields_}

    def kernel_info(self) -> dict:
        g, b, s, o = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _lib.check(_lib.lib().sabc_update_kernel_info(self._h, C.byref(g), C.byref(b), C.byref(s), C.byref(o)))
        return {"grid": g.value, "block": b.value, "smem_bytes": s.value, "blocks_per_sm": o.value}


class SABCstate:
    """Mirror of `mutable struct SABCstate` (src/SimulatedAnnealingABC.jl:28-42)."""

    def __init__(self, engine: Engine, algorithm: str):
        self._engine = engine
        self.algorithm = algorithm
        self.refresh()

    def refresh(self):
        eps, cnt = self._engine.get_state()
        self.eps = eps
        self.n_simulation, self.n_accept, self.n_resampling, self.n_population_updates = (int(c) for c in cnt)
        e, u, r = self._engine.get_history()
        self.eps_history = [row.copy() for row in e]
        self.u_history = [row.copy() for row in u]
        self.rho_history = [row.copy() for row in r]

    def cdfs_dist_prior(self, rho):
        """G of Albert et al. (2015): per-statistic ECDF transform, evaluated on the device."""
        rho = np.atleast_1d(np.asarray(rho, dtype=np.float64))
        out = np.empty_like(rho)
        for j in range(rho.size):
            k = self._engine.get_ecdf(j)
            x = np.array([rho[j]]); y = np.zeros(1)
            _lib.check(_lib.lib().sabc_ecdf_transform(_lib.ptr(k), k.size, _lib.ptr(x), 1, _lib.ptr(y)))
            out[j] = y[0]
        return out


# Julia field names
SABCstate.ϵ = property(lambda self: self.eps)
SABCstate.ϵ_history = property(lambda self: self.eps_history)
SABCstate.ρ_history = property(lambda self: self.rho_history)


class SABCresult:
    """Mirror of `struct SABCresult` (:55-60): population, u, ρ, state.  The arrays live on the device between calls and are
    fetched on first access; assigning to them marks the host copy as authoritative for the next update."""

    def __init__(self, engine: Engine, algorithm: str):
        self._engine = engine
        self.state = SABCstate(engine, algorithm)
        self._cache = None
        self._dirty = False

    def _fetch(self):
        if self._cache is None:
            self._cache = list(self._engine.get_population())
            for a in self._cache:
                a.flags.writeable = False      # an in-place edit would never reach the device: assign a whole array to the field instead
        return self._cache

    @property
    def population(self):
        th = self._fetch()[0]
        return th[:, 0] if th.shape[1] == 1 else th

    @population.setter
    def population(self, value):
        self._fetch()[0] = np.asarray(value, dtype=np.float64).reshape(self._engine.n_local, self._engine.d, order="F")
        self._dirty = True

    @property
    def u(self):
        return self._fetch()[1]

    @u.setter
    def u(self, value):
        self._fetch()[1] = np.asarray(value, dtype=np.float64).reshape(self._engine.n_local, self._engine.s, order="F")
        self._dirty = True

    @property
    def rho(self):
        return self._fetch()[2]

    @rho.setter
    def rho(self, value):
        self._fetch()[2] = np.asarray(value, dtype=np.float64).reshape(self._engine.n_local, self._engine.s, order="F")
        self._dirty = True

    def _push_if_dirty(self):
        if self._dirty:
            th, u, r = self._cache
            st = self.state
            self._engine.set_population(th, u, r, st.eps, [st.n_simulation, st.n_accept, st.n_resampling, st.n_population_updates])
            self._dirty = False

    def __len__(self):
        return self._engine.n_local

    def __repr__(self):   # show(io, ::SABCresult)  :65-82
        st = self.state
        n = self._engine.N
        mean_u = float(np.mean(self.u))
        denom = st.n_simulation - n
        acc = st.n_accept / denom if denom > 0 else float("nan")
        return (f"Approximate posterior sample with {n} particles:\n"
                f"  - algorithm: :{st.algorithm}\n"
                f"  - simulations used: {st.n_simulation}\n"
                f"  - number of population updates: {st.n_population_updates}\n"
                f"  - average transformed distance: {mean_u:.4g}\n"
                f"  - ϵ: {np.array2string(st.eps, precision=4)}\n"
                f"  - number of population resamplings: {st.n_resampling}\n"
                f"  - acceptance rate: {acc:.4g}\n"
                "The sample can be accessed with the field `population`.\n"
                "The history of ϵ can be accessed with the field `state.ϵ_history`.\n"
                "The history of ρ can be accessed with the field `state.ρ_history`.\n"
                "The history of u can be accessed with the field `state.u_history`.")


SABCresult.ρ = property(lambda self: self.rho)


def _resolve_model(f_dist, args, kwargs) -> DeviceModel:
    """`f_dist(θ, args...; kwargs...)` of the reference (src/SimulatedAnnealingABC.jl:163,174,315,421): the extra arguments are the data
    the closure works on.  On the device they are the model's parameter blob, so `f_dist` is either a DeviceModel (data already bound,
    no extra arguments) or a device-model FACTORY (`sabc_b200.models.gauss_mean`, ...) that the extra arguments are passed to:
    `sabc(sb.models.gauss_mean, prior, 1.0; sigma=1.0)`.  Arbitrary closures stay on the reference's CPU path."""
    if isinstance(f_dist, DeviceModel):
        if args or kwargs:
            raise TypeError("this DeviceModel already holds its data: pass the factory (e.g. sabc_b200.models.gauss_mean) to have extra "
                            "f_dist arguments bound into the parameter blob")
        return f_dist
    if callable(f_dist) and getattr(f_dist, "__module__", "").endswith(".models"):
        m = f_dist(*args, **kwargs)
        if isinstance(m, DeviceModel):
            return m
    raise TypeError("f_dist must be a DeviceModel or a device-model factory (sabc_b200.models.*): arbitrary closures stay on the "
                    "reference's CPU path (SimulatedAnnealingABC.jl); this package has no CPU fallback")


def _distributed_setup(comm):
    """comm = None (single GPU) | 'torch' (use torch.distributed's default group) | (rank, world, unique_id_bytes)."""
    if comm is None:
        return 0, 1, None
    if comm == "torch":
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        if world == 1:
            return 0, 1, None
        buf = C.create_string_buffer(128)
        if rank == 0:
            _lib.check(_lib.lib().sabc_nccl_unique_id(C.cast(buf, C.c_void_p)))
        obj = [bytes(buf.raw)]
        dist.broadcast_object_list(obj, src=0)
        return rank, world, obj[0]
    return comm


def sabc(f_dist, prior: Distribution, *args, n_particles: int = 100, n_simulation: int = 10_000,
         algorithm: str = "single_eps", proposal: Proposal | None = None, resample: int | None = None,
         v: float = 1.0, delta: float = 0.1, checkpoint_history: int = 1, show_progressbar: bool = False,
         show_checkpoint=math.inf, type: str | None = None, seed: int = 0x5ABC, device: int = -1, comm=None,
         flags: int = 0, ecdf_max_knots: int = 0, n_gpus: int = 0, gpu_ids=None, **kwargs) -> SABCresult:
    """sabc(f_dist, prior, args...; kw...)  -- src/SimulatedAnnealingABC.jl:451-492."""
    if "δ" in kwargs:
        delta = kwargs.pop("δ")
    if type is not None:
        if type not in TYPES:
            raise RuntimeError(f"Argument `type` must be :single, :multi or :hybrid, not `{type}`!")
        algorithm = TYPES[type]
    algorithm = str(algorithm).lstrip(":")
    if algorithm not in ALGORITHMS:                                                   # :462-464
        raise RuntimeError(f"Argument `algorithm` must be :multi_eps or :single_eps, not `{algorithm}`!")
    f_dist = _resolve_model(f_dist, args, kwargs)
    if n_simulation < n_particles:                                                    # :155-156
        raise RuntimeError(f"`n_simulation = {n_simulation}` is too small for {n_particles} particles.")
    if proposal is None:
        proposal = DifferentialEvolution(n_para=len(prior))                          # :454
    if resample is None:
        resample = 2 * n_particles                                                    # :455
    rank, world, uid = _distributed_setup(comm)
    eng = Engine(f_dist, prior, n_particles=n_particles, algorithm=algorithm, proposal=proposal, resample=resample,
                 v=v, delta=delta, seed=seed, device=device, rank=rank, world_size=world, nccl_unique_id=uid, flags=flags,
                 ecdf_max_knots=ecdf_max_knots, n_gpus=n_gpus, gpu_ids=gpu_ids)
    eng.init()                                                                        # :470-473
    res = SABCresult(eng, algorithm)
    n_sim_remaining = n_simulation - res.state.n_simulation                           # :478
    if n_sim_remaining < n_particles:
        warnings.warn("`n_simulation` too small to update all particles!")            # :479
    update_population(res, f_dist, prior, n_simulation=n_sim_remaining, resample=resample, proposal=proposal, v=v,
                      delta=delta, checkpoint_history=checkpoint_history)
    return res


def update_population(population_state: SABCresult, f_dist, prior: Distribution, *args, n_simulation: int, v: float = 1.0,
                      delta: float = 0.1, proposal: Proposal | None = None, resample: int | None = None,
                      checkpoint_history: int = 1, show_progressbar: bool = False, show_checkpoint=math.inf, **kwargs) -> SABCresult:
    """update_population!(population_state, f_dist, prior, args...; kw...)  -- src/SimulatedAnnealingABC.jl:251-402.
    Mutates and returns `population_state`."""
    if "δ" in kwargs:
        delta = kwargs.pop("δ")
    f_dist = _resolve_model(f_dist, args, kwargs)
    if v <= 0:
        raise RuntimeError("Annealing speed `v` must be positive.")                   # :261
    if delta <= 0:
        raise RuntimeError("Resamping intensity `δ` must be positive.")               # :262
    eng = population_state._engine
    if proposal is None:
        proposal = DifferentialEvolution(n_para=len(prior))                          # :255
    if resample is None:
        resample = 2 * eng.N                                                          # :256
    eng.set_tuning(v, delta, resample, proposal)
    population_state._push_if_dirty()
    eng.update(n_simulation, checkpoint_history)
    population_state._cache = None
    population_state.state.refresh()
    return population_state
