"""ctypes bindings of the two C ABIs of the engine.

* ``libiss_cuda.so``  -- include/iss_cuda.h, the CUDA hot path (yields, multiplicities, sampler,
  decays, QA).
* ``libiSS.so``       -- include/iss_host.h, the C forwarding layer over the drop-in C++ facade
  ``class iSS`` (same public API as reference src/iSS.h:16-102).

This module is plumbing for tests and bench.py.  It has no fallback: if the shared libraries are
missing it raises, it never computes anything itself.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(_HERE)
TABLES = os.path.join(REPO, "iSS_tables")

NFIELD = 28
F = {n: i for i, n in enumerate(
    "tau x y eta da0 da1 da2 da3 ut ux uy uz e T P nB muB muS muQ bulkPi "
    "pixx pixy pixz piyy piyz qx qy qz".split())}
QA_NSPEC, QA_NPT, QA_NY, QA_NPHI, QA_NV2, QA_HEAD = 16, 100, 100, 64, 20, 32
QA_PER = 3*QA_NPT + QA_NY + QA_NPHI + 2*QA_NV2 + 2
T_KINDS = ["yields", "scan", "mult", "sample", "decay", "qa", "setup"]

TABLE_BESSEL_K, TABLE_EXPINT, TABLE_CE, TABLE_MOM22, TABLE_MOM14, TABLE_KAPPA_B, TABLE_BULK14 = 1, 2, 3, 4, 5, 6, 7


class Species(C.Structure):
    _fields_ = [("pid", C.c_int32), ("gspin", C.c_int32), ("baryon", C.c_int32),
                ("strange", C.c_int32), ("charge", C.c_int32), ("sign", C.c_int32),
                ("decay_idx", C.c_int32), ("reserved", C.c_int32), ("mass", C.c_double)]


class Options(C.Structure):
    _fields_ = [("hydro_mode", C.c_int32), ("include_deltaf_shear", C.c_int32),
                ("include_deltaf_bulk", C.c_int32), ("include_deltaf_diffusion", C.c_int32),
                ("bulk_deltaf_kind", C.c_int32), ("dN_dy_sampling_model", C.c_int32),
                ("local_charge_conservation", C.c_int32), ("reserved", C.c_int32),
                ("dN_dy_sampling_para1", C.c_double), ("y_LB", C.c_double), ("y_RB", C.c_double)]


class DecaySpecies(C.Structure):
    _fields_ = [("pid", C.c_int32), ("stable", C.c_int32), ("n_channels", C.c_int32),
                ("first_channel", C.c_int32), ("baryon", C.c_int32), ("strange", C.c_int32),
                ("charge", C.c_int32), ("reserved", C.c_int32), ("mass", C.c_double),
                ("width", C.c_double)]


class DecayChannel(C.Structure):
    _fields_ = [("n_part", C.c_int32), ("daughter", C.c_int32*5),
                ("branching_ratio", C.c_double)]


class SpectraOptions(C.Structure):
    _fields_ = [("include_deltaf_shear", C.c_int32), ("include_deltaf_bulk", C.c_int32),
                ("bulk_deltaf_kind", C.c_int32), ("include_deltaf_diffusion", C.c_int32),
                ("restrict_deltaf", C.c_int32), ("use_pos_dN_only", C.c_int32),
                ("deltaf_max_ratio", C.c_double)]


class LegacyOptions(C.Structure):
    _fields_ = [("include_deltaf_shear", C.c_int32), ("include_deltaf_bulk", C.c_int32),
                ("bulk_deltaf_kind", C.c_int32), ("include_deltaf_diffusion", C.c_int32),
                ("restrict_deltaf", C.c_int32), ("reserved", C.c_int32),
                ("deltaf_max_ratio", C.c_double), ("sample_pT_up_to", C.c_double),
                ("sample_y_minus_eta_s_range", C.c_double)]


class IngestOptions(C.Structure):
    _fields_ = [("boost_invariant", C.c_int32), ("regulate_eos", C.c_int32), ("hrg_nB", C.c_int32),
                ("reserved", C.c_int32), ("hrg_rows", C.c_int64)]


class IngestResult(C.Structure):
    _fields_ = [("n_in", C.c_int64), ("n_after_T", C.c_int64), ("n_kept", C.c_int64)]


class Counts(C.Structure):
    _fields_ = [("n_events", C.c_int64), ("n_hadrons", C.c_int64), ("n_tries", C.c_int64),
                ("n_cell_redraws", C.c_int64)]


HADRON_DTYPE = np.dtype([("pid", "<i4"), ("mass", "<f4"), ("E", "<f4"), ("px", "<f4"),
                         ("py", "<f4"), ("pz", "<f4"), ("t", "<f4"), ("x", "<f4"), ("y", "<f4"),
                         ("z", "<f4")])
assert HADRON_DTYPE.itemsize == 40

SPECIES_DTYPE = np.dtype([("pid", "<i4"), ("gspin", "<i4"), ("baryon", "<i4"), ("strange", "<i4"),
                          ("charge", "<i4"), ("sign", "<i4"), ("decay_idx", "<i4"),
                          ("reserved", "<i4"), ("mass", "<f8")])
assert SPECIES_DTYPE.itemsize == C.sizeof(Species)

# every symbol include/iss_cuda.h declares (tests check that the library exports them all)
CUDA_SYMBOLS = [
    "iss_cuda_create", "iss_cuda_destroy", "iss_cuda_last_error", "iss_cuda_set_stream",
    "iss_cuda_synchronize", "iss_cuda_upload_surface", "iss_cuda_upload_species",
    "iss_cuda_upload_table", "iss_cuda_upload_decay_table", "iss_cuda_set_options",
    "iss_cuda_compute_yields", "iss_cuda_sample", "iss_cuda_get_multiplicities",
    "iss_cuda_get_poisson_params", "iss_cuda_decay", "iss_cuda_event_offsets",
    "iss_cuda_fetch_event", "iss_cuda_fetch_all", "iss_cuda_device_hadrons", "iss_cuda_qa_size",
    "iss_cuda_histograms", "iss_cuda_qa_device_ptr", "iss_cuda_qa_fetch", "iss_cuda_timing",
    "iss_cuda_mem_info", "iss_cuda_host_alloc", "iss_cuda_host_free", "iss_cuda_fp64_peak",
    "iss_cuda_set_trace", "iss_cuda_get_trace", "iss_cuda_upload_surface_aos",
    "iss_cuda_fetch_all_async", "iss_cuda_fetch_wait", "iss_cuda_sample_momentum",
    "iss_cuda_upload_surface_lab", "iss_cuda_spectra", "iss_cuda_spectra_stats",
    "iss_cuda_ingest_music_binary", "iss_cuda_set_surface_chunk", "iss_cuda_chunk_yields_local",
    "iss_cuda_chunk_yields_finish", "iss_cuda_chunk_yields_allgather", "iss_cuda_chunk_block_yields", "iss_cuda_upload_surface_aos_part",
    "iss_cuda_legacy_upload_positions", "iss_cuda_legacy_upload_z_table",
    "iss_cuda_legacy_set_options", "iss_cuda_legacy_compute_yields",
    "iss_cuda_nccl_unique_id", "iss_cuda_nccl_init", "iss_cuda_nccl_finalize",
    "iss_cuda_histograms_allreduce",
]
HOST_SYMBOLS = [
    "iss_host_create", "iss_host_destroy", "iss_host_set_param", "iss_host_get_param",
    "iss_host_parse_param", "iss_host_set_random_seed", "iss_host_read_in_FO_surface",
    "iss_host_generate_samples", "iss_host_shell", "iss_host_perform_checks",
    "iss_host_get_number_of_sampled_events", "iss_host_get_number_of_particles",
    "iss_host_get_hadron_list_iev", "iss_host_clear", "iss_host_prepare_sampler",
    "iss_host_cuda_handle", "iss_host_lrf_surface", "iss_host_species", "iss_host_hadron_buffer",
    "iss_host_species_dN", "iss_host_qa_block", "iss_host_write_samples", "iss_host_spectra_table",
]

_cuda = None
_host = None


def cuda_lib_path():
    return os.path.join(_HERE, "libiss_cuda.so")


def host_lib_path():
    return os.path.join(_HERE, "libiSS.so")


def cuda_lib():
    """libiss_cuda.so with argument types declared.  Raises OSError if it has not been built."""
    global _cuda
    if _cuda is not None:
        return _cuda
    L = C.CDLL(cuda_lib_path(), mode=C.RTLD_GLOBAL)
    vp, i32, i64, u64, dp = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.POINTER(C.c_double)
    i64p = C.POINTER(C.c_int64)
    sig = {
        "iss_cuda_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "iss_cuda_destroy": (C.c_int, [vp]),
        "iss_cuda_last_error": (C.c_char_p, [vp]),
        "iss_cuda_set_stream": (C.c_int, [vp, vp]),
        "iss_cuda_synchronize": (C.c_int, [vp]),
        "iss_cuda_upload_surface": (C.c_int, [vp, C.POINTER(vp), i64]),
        "iss_cuda_upload_species": (C.c_int, [vp, vp, i32]),
        "iss_cuda_upload_table": (C.c_int, [vp, i32, vp, i64, i64, vp]),
        "iss_cuda_upload_decay_table": (C.c_int, [vp, vp, i32, vp, i32]),
        "iss_cuda_set_options": (C.c_int, [vp, C.POINTER(Options)]),
        "iss_cuda_compute_yields": (C.c_int, [vp, vp, vp]),
        "iss_cuda_set_surface_chunk": (C.c_int, [vp, i64, i64]),
        "iss_cuda_chunk_yields_local": (C.c_int, [vp, C.POINTER(vp), i64p]),
        "iss_cuda_chunk_yields_finish": (C.c_int, [vp, C.POINTER(vp), i64p, i32, C.c_int, vp]),
        "iss_cuda_chunk_yields_allgather": (C.c_int, [vp, i64p, i32, vp, vp]),
        "iss_cuda_chunk_block_yields": (C.c_int, [vp, vp, i64]),
        "iss_cuda_sample": (C.c_int, [vp, u64, i64, i64, C.POINTER(Counts)]),
        "iss_cuda_get_multiplicities": (C.c_int, [vp, vp]),
        "iss_cuda_get_poisson_params": (C.c_int, [vp, vp, vp]),
        "iss_cuda_decay": (C.c_int, [vp, u64, C.POINTER(Counts)]),
        "iss_cuda_event_offsets": (C.c_int, [vp, vp]),
        "iss_cuda_fetch_event": (C.c_int, [vp, i64, vp, i64, i64p]),
        "iss_cuda_fetch_all": (C.c_int, [vp, vp, i64, i64p]),
        "iss_cuda_device_hadrons": (C.c_int, [vp, C.POINTER(vp), i64p]),
        "iss_cuda_qa_size": (i64, []),
        "iss_cuda_histograms": (C.c_int, [vp, vp, i32, C.c_int]),
        "iss_cuda_qa_device_ptr": (C.c_int, [vp, C.POINTER(vp)]),
        "iss_cuda_qa_fetch": (C.c_int, [vp, vp]),
        "iss_cuda_timing": (C.c_int, [vp, C.c_int, vp, vp, C.c_int]),
        "iss_cuda_mem_info": (C.c_int, [vp, i64p, i64p]),
        "iss_cuda_host_alloc": (C.c_int, [vp, C.POINTER(vp), i64]),
        "iss_cuda_host_free": (C.c_int, [vp, vp]),
        "iss_cuda_fp64_peak": (C.c_int, [vp, dp]),
        "iss_cuda_upload_surface_aos": (C.c_int, [vp, vp, i64]),
        "iss_cuda_upload_surface_aos_part": (C.c_int, [vp, vp, i64, i64, i64]),
        "iss_cuda_fetch_all_async": (C.c_int, [vp, vp, i64, i64p]),
        "iss_cuda_fetch_wait": (C.c_int, [vp]),
        "iss_cuda_sample_momentum": (C.c_int, [vp, C.c_double, C.c_double, C.c_double, i32, i64, u64, vp]),
        "iss_cuda_set_trace": (C.c_int, [vp, C.c_int]),
        "iss_cuda_get_trace": (C.c_int, [vp, vp, vp]),
        "iss_cuda_upload_surface_lab": (C.c_int, [vp, vp, i64]),
        "iss_cuda_spectra": (C.c_int, [vp, C.POINTER(SpectraOptions), vp, i32, vp, i32, vp, i32, vp, vp,
                                       i32, vp, vp]),
        "iss_cuda_spectra_stats": (C.c_int, [vp, dp, dp]),
        "iss_cuda_legacy_upload_positions": (C.c_int, [vp, vp, i64]),
        "iss_cuda_legacy_upload_z_table": (C.c_int, [vp, vp, vp, i32]),
        "iss_cuda_legacy_set_options": (C.c_int, [vp, C.POINTER(LegacyOptions)]),
        "iss_cuda_legacy_compute_yields": (C.c_int, [vp, vp, vp, vp]),
        "iss_cuda_ingest_music_binary": (C.c_int, [vp, vp, i64, C.POINTER(IngestOptions), vp, vp, vp, vp,
                                                   C.POINTER(IngestResult)]),
        "iss_cuda_nccl_unique_id": (C.c_int, [vp]),
        "iss_cuda_nccl_init": (C.c_int, [vp, vp, i32, i32]),
        "iss_cuda_nccl_finalize": (C.c_int, [vp]),
        "iss_cuda_histograms_allreduce": (C.c_int, [vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _cuda = L
    return L


def host_lib():
    """libiSS.so (C layer over class iSS).  Raises OSError if it has not been built."""
    global _host
    if _host is not None:
        return _host
    cuda_lib()
    L = C.CDLL(host_lib_path(), mode=C.RTLD_GLOBAL)
    vp, i64p = C.c_void_p, C.POINTER(C.c_int64)
    cs = C.c_char_p
    sig = {
        "iss_host_create": (vp, [cs, cs, cs, cs, cs]),
        "iss_host_destroy": (None, [vp]),
        "iss_host_set_param": (None, [vp, cs, C.c_double]),
        "iss_host_get_param": (C.c_double, [vp, cs, C.c_double]),
        "iss_host_parse_param": (None, [vp, cs]),
        "iss_host_set_random_seed": (None, [vp, C.c_int]),
        "iss_host_read_in_FO_surface": (C.c_int, [vp]),
        "iss_host_generate_samples": (C.c_int, [vp]),
        "iss_host_shell": (C.c_int, [vp]),
        "iss_host_perform_checks": (None, [vp]),
        "iss_host_get_number_of_sampled_events": (C.c_int, [vp]),
        "iss_host_get_number_of_particles": (C.c_int, [vp, C.c_int]),
        "iss_host_get_hadron_list_iev": (vp, [vp, C.c_int, i64p]),
        "iss_host_clear": (None, [vp]),
        "iss_host_prepare_sampler": (C.c_int, [vp]),
        "iss_host_cuda_handle": (vp, [vp]),
        "iss_host_lrf_surface": (C.c_int64, [vp, vp]),
        "iss_host_species": (C.c_int32, [vp, vp]),
        "iss_host_hadron_buffer": (vp, [vp, C.POINTER(vp), i64p]),
        "iss_host_species_dN": (C.c_int32, [vp, vp]),
        "iss_host_qa_block": (C.c_int, [vp, vp]),
        "iss_host_write_samples": (C.c_int, [C.c_int, vp, vp, C.c_int64, cs]),
        "iss_host_spectra_table": (C.c_int, [vp, C.c_int32, vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                             C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _host = L
    return L


def write_samples(fmt, hadrons, event_offsets, table_path=TABLES):
    """reference-format sample files in the current directory; fmt: "oscar", "gzip", "binary"."""
    h = np.ascontiguousarray(hadrons, dtype=HADRON_DTYPE)
    off = np.ascontiguousarray(event_offsets, dtype=np.int64)
    code = {"oscar": 0, "gzip": 1, "binary": 2}[fmt]
    rc = host_lib().iss_host_write_samples(code, h.ctypes.data_as(C.c_void_p),
                                           off.ctypes.data_as(C.c_void_p), len(off) - 1,
                                           table_path.encode())
    if rc != 0:
        raise IssError("iss_host_write_samples failed")


class IssError(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Engine:
    """Thin object wrapper of one ``iss_handle`` (one GPU)."""

    def __init__(self, device=0, handle=None):
        self.L = cuda_lib()
        self.owned = handle is None
        if handle is None:
            h = C.c_void_p()
            rc = self.L.iss_cuda_create(device, C.byref(h))
            if rc != 0:
                raise IssError("iss_cuda_create failed (status %d): no usable CUDA device; the "
                               "engine has no CPU fallback" % rc)
            self.h = h
        else:
            self.h = C.c_void_p(handle)
        self.nspecies = 0
        self.ncell = 0

    def close(self):
        if self.owned and self.h:
            self.L.iss_cuda_destroy(self.h)
        self.h = None

    def check(self, rc, what):
        if rc != 0:
            raise IssError("%s failed (status %d): %s" %
                           (what, rc, self.L.iss_cuda_last_error(self.h).decode()))

    # ---- inputs
    def upload_surface(self, cells):
        """cells: float32 [ncell, 28] in ISS_F_* order (AoS); uploaded as SoA."""
        cells = np.ascontiguousarray(cells, dtype=np.float32)
        soa = np.ascontiguousarray(cells.T)
        ptrs = (C.c_void_p*NFIELD)(*[soa[k].ctypes.data for k in range(NFIELD)])
        self.check(self.L.iss_cuda_upload_surface(self.h, ptrs, cells.shape[0]), "upload_surface")
        self.ncell = cells.shape[0]

    def upload_surface_parts(self, cells, nparts):
        """the same records through iss_cuda_upload_surface_aos_part, in `nparts` pieces"""
        cells = np.ascontiguousarray(cells, dtype=np.float32)
        n = cells.shape[0]
        for k in range(nparts):
            b, e = n*k//nparts, n*(k + 1)//nparts
            if e > b:
                self.check(self.L.iss_cuda_upload_surface_aos_part(self.h, _ptr(cells[b:e]), b, e - b, n),
                           "upload_surface_aos_part")
        self.ncell = n

    def upload_species(self, species):
        """species: structured array of SPECIES_DTYPE."""
        species = np.ascontiguousarray(species, dtype=SPECIES_DTYPE)
        self.check(self.L.iss_cuda_upload_species(self.h, _ptr(species), len(species)),
                   "upload_species")
        self.nspecies = len(species)

    def upload_table(self, kind, data, n0, n1=0, grid=None):
        data = np.ascontiguousarray(data, dtype=np.float64)
        g = None if grid is None else _ptr(np.ascontiguousarray(grid, dtype=np.float64))
        self.check(self.L.iss_cuda_upload_table(self.h, kind, _ptr(data), n0, n1, g),
                   "upload_table")

    def set_options(self, **kw):
        o = Options()
        o.hydro_mode = 2
        o.dN_dy_sampling_model = 30
        o.y_LB, o.y_RB = -5.0, 5.0
        for k, v in kw.items():
            setattr(o, k, v)
        self.check(self.L.iss_cuda_set_options(self.h, C.byref(o)), "set_options")

    # ---- hot path
    def compute_yields(self, want_cells=False):
        dN = np.zeros(self.nspecies)
        y = np.zeros((self.nspecies, self.ncell)) if want_cells else None
        self.check(self.L.iss_cuda_compute_yields(self.h, _ptr(dN), _ptr(y) if want_cells else None),
                   "compute_yields")
        return (dN, y) if want_cells else dN

    # ---- surface-chunk sharding (include/iss_cuda.h): this handle holds a contiguous cell range
    def set_surface_chunk(self, cell_begin, ncell_global):
        self.check(self.L.iss_cuda_set_surface_chunk(self.h, cell_begin, ncell_global),
                   "set_surface_chunk")

    def chunk_yields_local(self):
        """-> (device pointer of the [nspecies][ntile_local] tile sums, ntile_local)"""
        p = C.c_void_p()
        n = C.c_int64()
        self.check(self.L.iss_cuda_chunk_yields_local(self.h, C.byref(p), C.byref(n)),
                   "chunk_yields_local")
        return p.value, n.value

    def chunk_tilesums_host(self):
        """the same block copied to the host (tests; the multi-GPU path all-gathers on the device)"""
        import torch
        ptr, nt = self.chunk_yields_local()
        from .sharding import device_block_as_tensor
        t = device_block_as_tensor(ptr, self.nspecies*nt, torch.device("cuda", torch.cuda.current_device()))
        return t.cpu().numpy().reshape(self.nspecies, nt).copy()

    def chunk_yields_finish(self, blocks, ntiles, on_device):
        """blocks: per rank, a device pointer (int) or a host float64 array [nspecies][ntiles[r]]"""
        nr = len(blocks)
        keep = []
        ptrs = (C.c_void_p*nr)()
        for r, b in enumerate(blocks):
            if on_device:
                ptrs[r] = int(b)
            else:
                a = np.ascontiguousarray(b, dtype=np.float64)
                keep.append(a)
                ptrs[r] = a.ctypes.data
        nt = (C.c_int64*nr)(*[int(x) for x in ntiles])
        dN = np.zeros(self.nspecies)
        self.check(self.L.iss_cuda_chunk_yields_finish(self.h, ptrs, nt, nr, 1 if on_device else 0,
                                                       _ptr(dN)), "chunk_yields_finish")
        return dN

    def chunk_block_yields(self, ncell_global):
        """yield (all species) of every 4096-cell block of the whole surface, after the finish step"""
        nb = (int(ncell_global) + 4095)//4096
        out = np.zeros(nb)
        self.check(self.L.iss_cuda_chunk_block_yields(self.h, _ptr(out), nb), "chunk_block_yields")
        return out

    def chunk_yields_allgather(self, ntiles, comm=None):
        """local yields, all-gather of the tile sums over the handle's (or the given) NCCL communicator
        and the global part, one call on the handle's stream; ntiles: tiles of every rank's chunk"""
        nr = len(ntiles)
        nt = (C.c_int64*nr)(*[int(x) for x in ntiles])
        dN = np.zeros(self.nspecies)
        self.check(self.L.iss_cuda_chunk_yields_allgather(self.h, nt, nr, comm, _ptr(dN)),
                   "chunk_yields_allgather")
        return dN

    def sample(self, seed, ev_begin, ev_end):
        c = Counts()
        self.check(self.L.iss_cuda_sample(self.h, seed, ev_begin, ev_end, C.byref(c)), "sample")
        return c

    def decay(self, seed):
        c = Counts()
        self.check(self.L.iss_cuda_decay(self.h, seed, C.byref(c)), "decay")
        return c

    def multiplicities(self, nev):
        m = np.zeros((nev, self.nspecies), dtype=np.int64)
        self.check(self.L.iss_cuda_get_multiplicities(self.h, _ptr(m)), "get_multiplicities")
        return m

    def poisson_params(self):
        lam = np.zeros(self.nspecies)
        pm = np.zeros(self.nspecies)
        self.check(self.L.iss_cuda_get_poisson_params(self.h, _ptr(lam), _ptr(pm)),
                   "get_poisson_params")
        return lam, pm

    def event_offsets(self, nev):
        off = np.zeros(nev + 1, dtype=np.int64)
        self.check(self.L.iss_cuda_event_offsets(self.h, _ptr(off)), "event_offsets")
        return off

    def fetch_all(self):
        n = C.c_int64()
        self.check(self.L.iss_cuda_fetch_all(self.h, None, 0, C.byref(n)), "fetch_all(size)")
        out = np.zeros(n.value, dtype=HADRON_DTYPE)
        if n.value:
            self.check(self.L.iss_cuda_fetch_all(self.h, _ptr(out), n.value, C.byref(n)),
                       "fetch_all")
        return out

    def sample_momentum(self, mass, T, mu, sign, n, seed):
        out = np.zeros(n)
        self.check(self.L.iss_cuda_sample_momentum(self.h, mass, T, mu, sign, n, seed, _ptr(out)),
                   "sample_momentum")
        return out

    def set_trace(self, enable):
        self.check(self.L.iss_cuda_set_trace(self.h, int(enable)), "set_trace")

    def get_trace(self, n):
        cell = np.zeros(n, dtype=np.int32)
        tries = np.zeros(n, dtype=np.int32)
        self.check(self.L.iss_cuda_get_trace(self.h, _ptr(cell), _ptr(tries)), "get_trace")
        return cell, tries

    def histograms(self, pids, accumulate=False):
        p = np.ascontiguousarray(pids, dtype=np.int32)
        self.check(self.L.iss_cuda_histograms(self.h, _ptr(p), len(p), int(accumulate)),
                   "histograms")
        qa = np.zeros(self.L.iss_cuda_qa_size())
        self.check(self.L.iss_cuda_qa_fetch(self.h, _ptr(qa)), "qa_fetch")
        return qa

    def qa_device_ptr(self):
        p = C.c_void_p()
        self.check(self.L.iss_cuda_qa_device_ptr(self.h, C.byref(p)), "qa_device_ptr")
        return p.value

    def timing(self, enable=True, reset=False):
        ms = np.zeros(len(T_KINDS))
        n = np.zeros(len(T_KINDS), dtype=np.int64)
        self.check(self.L.iss_cuda_timing(self.h, int(enable), _ptr(ms), _ptr(n), int(reset)),
                   "timing")
        return dict(zip(T_KINDS, ms)), dict(zip(T_KINDS, n))

    # ---- smooth Cooper-Frye spectra (EmissionFunctionArray::calculate_dN_pTdpTdphidy)
    def upload_surface_lab(self, cells):
        """cells: float32 [ncell, 32] lab-frame records in ISS_L_* order."""
        cells = np.ascontiguousarray(cells, dtype=np.float32)
        assert cells.ndim == 2 and cells.shape[1] == 32
        self.check(self.L.iss_cuda_upload_surface_lab(self.h, _ptr(cells), cells.shape[0]),
                   "upload_surface_lab")

    def spectra(self, species, pT, phi, y_minus_eta, y_weight, **opt):
        """Returns (dN, dN_max), each [nspecies, npT, nphi]."""
        species = np.ascontiguousarray(species, dtype=SPECIES_DTYPE)
        pT = np.ascontiguousarray(pT, dtype=np.float64)
        phi = np.ascontiguousarray(phi, dtype=np.float64)
        y = np.ascontiguousarray(y_minus_eta, dtype=np.float64)
        w = np.ascontiguousarray(y_weight, dtype=np.float64)
        o = SpectraOptions()
        o.include_deltaf_shear = 1
        o.restrict_deltaf = 1
        o.deltaf_max_ratio = 1.0
        o.bulk_deltaf_kind = 1
        for k, v in opt.items():
            setattr(o, k, v)
        dN = np.zeros((len(species), len(pT), len(phi)))
        dN_max = np.zeros_like(dN)
        self.check(self.L.iss_cuda_spectra(self.h, C.byref(o), _ptr(species), len(species), _ptr(pT),
                                           len(pT), _ptr(phi), len(phi), _ptr(y), _ptr(w), len(y),
                                           _ptr(dN), _ptr(dN_max)), "spectra")
        return dN, dN_max

    # ---- legacy "conventional" sampler (MC_sampling = 2, EmissionFunctionArray)
    def legacy_setup(self, lab, pos, ztab, **opt):
        """lab: float32 [ncell, 32] (ISS_L_* order); pos: float32 [ncell, 4] x, y, eta_s, 0;
        ztab: the two columns of iSS_tables/z_exp_m_z.dat; opt: fields of iss_legacy_options."""
        self.upload_surface_lab(lab)
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        assert pos.shape == (len(lab), 4)
        self.check(self.L.iss_cuda_legacy_upload_positions(self.h, _ptr(pos), len(pos)),
                   "legacy_upload_positions")
        zx = np.ascontiguousarray(ztab[0], dtype=np.float64)
        zy = np.ascontiguousarray(ztab[1], dtype=np.float64)
        self.check(self.L.iss_cuda_legacy_upload_z_table(self.h, _ptr(zx), _ptr(zy), len(zx)),
                   "legacy_upload_z_table")
        o = LegacyOptions()
        o.deltaf_max_ratio = 1.0
        o.bulk_deltaf_kind = 1
        for k, v in opt.items():
            setattr(o, k, v)
        self.check(self.L.iss_cuda_legacy_set_options(self.h, C.byref(o)), "legacy_set_options")
        self.ncell = len(lab)

    def legacy_compute_yields(self, want_cells=False, want_maximum=False):
        dN = np.zeros(self.nspecies)
        y = np.zeros((self.nspecies, self.ncell)) if want_cells else None
        mx = np.zeros((self.nspecies, self.ncell)) if want_maximum else None
        self.check(self.L.iss_cuda_legacy_compute_yields(
            self.h, _ptr(dN), _ptr(y) if want_cells else None, _ptr(mx) if want_maximum else None),
            "legacy_compute_yields")
        return dN, y, mx

    def spectra_stats(self):
        """(evaluations, kernel milliseconds) of the last spectra() call."""
        n, ms = C.c_double(), C.c_double()
        self.check(self.L.iss_cuda_spectra_stats(self.h, C.byref(n), C.byref(ms)), "spectra_stats")
        return n.value, ms.value

    # ---- surface ingest on the device (binary MUSIC records -> LRF records)
    def ingest_music_binary(self, raw, hrg=None, hrg_nB=1, boost_invariant=False):
        """raw: float32 [ncell, 34]; hrg: float64 [rows, 7] or None (no EOS regulation).
        Returns (lrf [n_kept, 28], tmunu [n_after_T, 16], status [ncell])."""
        raw = np.ascontiguousarray(raw, dtype=np.float32)
        n = raw.shape[0]
        o = IngestOptions()
        o.boost_invariant = int(bool(boost_invariant))
        o.regulate_eos = 0 if hrg is None else 1
        o.hrg_nB = int(hrg_nB)
        hp = None
        if hrg is not None:
            hrg = np.ascontiguousarray(hrg, dtype=np.float64)
            o.hrg_rows = hrg.shape[0]
            hp = _ptr(hrg)
        lrf = np.zeros((n, NFIELD), dtype=np.float32)
        tm = np.zeros((n, 16), dtype=np.float32)
        st = np.zeros(n, dtype=np.uint8)
        res = IngestResult()
        self.check(self.L.iss_cuda_ingest_music_binary(self.h, _ptr(raw), n, C.byref(o), hp, _ptr(lrf),
                                                       _ptr(tm), _ptr(st), C.byref(res)),
                   "ingest_music_binary")
        return lrf[:res.n_kept], tm[:res.n_after_T], st

    def fp64_peak(self):
        t = C.c_double()
        self.check(self.L.iss_cuda_fp64_peak(self.h, C.byref(t)), "fp64_peak")
        return t.value

    def set_stream(self, cuda_stream):
        self.check(self.L.iss_cuda_set_stream(self.h, C.c_void_p(cuda_stream)), "set_stream")

    def synchronize(self):
        self.check(self.L.iss_cuda_synchronize(self.h), "synchronize")


class Sampler:
    """Python face of the drop-in ``class iSS`` (reference src/iSS.h:16-102) via libiSS.so."""

    def __init__(self, path, param_file, surface_filename="surface.dat", table_path=TABLES,
                 particle_table_path=None, **overrides):
        self.L = host_lib()
        ptp = particle_table_path or table_path
        self.s = C.c_void_p(self.L.iss_host_create(path.encode(), table_path.encode(),
                                                   ptp.encode(), param_file.encode(),
                                                   surface_filename.encode()))
        for k, v in overrides.items():
            self.set_param(k, v)

    def close(self):
        if self.s:
            self.L.iss_host_destroy(self.s)
        self.s = None

    def set_param(self, name, value):
        self.L.iss_host_set_param(self.s, name.encode(), float(value))

    def get_param(self, name, default=0.0):
        return self.L.iss_host_get_param(self.s, name.encode(), default)

    def set_random_seed(self, seed):
        self.L.iss_host_set_random_seed(self.s, int(seed))

    def read_in_FO_surface(self):
        return self.L.iss_host_read_in_FO_surface(self.s)

    def generate_samples(self):
        return self.L.iss_host_generate_samples(self.s)

    def shell(self):
        return self.L.iss_host_shell(self.s)

    def perform_checks(self):
        self.L.iss_host_perform_checks(self.s)

    def get_number_of_sampled_events(self):
        return self.L.iss_host_get_number_of_sampled_events(self.s)

    def get_number_of_particles(self, iev):
        return self.L.iss_host_get_number_of_particles(self.s, iev)

    def get_hadron_list_iev(self, iev):
        n = C.c_int64()
        p = self.L.iss_host_get_hadron_list_iev(self.s, iev, C.byref(n))
        if n.value == 0:
            return np.zeros(0, dtype=HADRON_DTYPE)
        buf = (C.c_char*(40*n.value)).from_address(p)
        return np.frombuffer(buf, dtype=HADRON_DTYPE).copy()

    # ---- engine additions
    def prepare_sampler(self):
        return self.L.iss_host_prepare_sampler(self.s)

    def engine(self):
        h = self.L.iss_host_cuda_handle(self.s)
        if not h:
            raise IssError("sampler not prepared")
        e = Engine(handle=h)
        e.nspecies = self.L.iss_host_species(self.s, None)
        e.ncell = self.L.iss_host_lrf_surface(self.s, None)
        return e

    def lrf_surface(self):
        n = self.L.iss_host_lrf_surface(self.s, None)
        a = np.zeros((n, NFIELD), dtype=np.float32)
        self.L.iss_host_lrf_surface(self.s, _ptr(a))
        return a

    def species(self):
        n = self.L.iss_host_species(self.s, None)
        a = np.zeros(n, dtype=SPECIES_DTYPE)
        self.L.iss_host_species(self.s, _ptr(a))
        return a

    def species_dN(self):
        n = self.L.iss_host_species_dN(self.s, None)
        a = np.zeros(n)
        self.L.iss_host_species_dN(self.s, _ptr(a))
        return a

    def hadrons(self):
        """(all hadrons, event offsets) as numpy views of the sampler-owned pinned buffer."""
        off_p = C.c_void_p()
        nev = C.c_int64()
        p = self.L.iss_host_hadron_buffer(self.s, C.byref(off_p), C.byref(nev))
        off = np.frombuffer((C.c_int64*(nev.value + 1)).from_address(off_p.value),
                            dtype=np.int64).copy()
        n = int(off[-1])
        if n == 0:
            return np.zeros(0, dtype=HADRON_DTYPE), off
        buf = (C.c_char*(40*n)).from_address(p)
        return np.frombuffer(buf, dtype=HADRON_DTYPE), off

    def spectra_table(self, monval):
        """dN/(pT dpT dphi dy) [npT, nphi] of one species after generate_samples() with
        MC_sampling = 0, calculate_vn = 1; also returns (kernel_ms, evaluations) of the run."""
        npt, nphi, ms, ev = C.c_int32(), C.c_int32(), C.c_double(), C.c_double()
        if self.L.iss_host_spectra_table(self.s, monval, None, C.byref(npt), C.byref(nphi),
                                         C.byref(ms), C.byref(ev)) != 0:
            raise IssError("no spectra table for species %d" % monval)
        a = np.zeros((npt.value, nphi.value))
        self.L.iss_host_spectra_table(self.s, monval, _ptr(a), None, None, None, None)
        return a, ms.value, ev.value

    def qa_block(self):
        qa = np.zeros(cuda_lib().iss_cuda_qa_size())
        if self.L.iss_host_qa_block(self.s, _ptr(qa)) != 0:
            raise IssError("no QA block")
        return qa
