"""ctypes prototypes for every symbol of include/dbs_b200.h.

Loading fails loudly if lib/libdbs_b200.so is missing (build it with `python __graft_entry__.py`
or `make -C diffusion-boltzmann-sampler_b200/csrc`).  There is no fallback implementation.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DBS_B200_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libdbs_b200.so"))

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"libdbs_b200.so not found at {LIB_PATH}: the sm_100a engine is not built. "
        "Run `python __graft_entry__.py` (or `make -C diffusion-boltzmann-sampler_b200/csrc`). "
        "There is no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED, ERR_NO_DEVICE = 0, -1, -2, -3, -4

STREAM_SWEEP, STREAM_INIT, STREAM_NOISE = 0x53574550, 0x494E4954, 0x4E4F4953
STEP_X0 = -1

vp, i32, i64, u64, f32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float


class UnetConfig(C.Structure):
    _fields_ = [("in_channels", i32), ("base_channels", i32), ("time_embed_dim", i32), ("num_blocks", i32),
                ("output_scale", f32), ("normalize_output", i32)]


class CsnConfig(C.Structure):
    _fields_ = [("n_particles", i32), ("hidden_dim", i32), ("time_embed_dim", i32), ("pos_embed_dim", i32),
                ("num_blocks", i32), ("num_heads", i32), ("use_attention", i32)]


class Param(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", vp), ("numel", i64), ("is_device", i32)]


# name -> (restype, argtypes); kept in one table so tests can check the library exports all of them
PROTOTYPES = {
    "dbs_last_error": (C.c_char_p, []),
    "dbs_abi_version": (i32, []),
    "dbs_built_sm": (i32, []),
    "dbs_launch_count": (u64, []),
    "dbs_profile_enable": (i32, [i32]),
    "dbs_profile_num_classes": (i32, []),
    "dbs_profile_class_name": (C.c_char_p, [i32]),
    "dbs_profile_read": (i32, [vp, vp, i32, i32]),
    "dbs_profile_dump": (i32, [C.c_char_p]),
    "dbs_ising_words_per_chain": (i32, [i32]),
    "dbs_ising_init_packed": (i32, [vp, i32, i64, u64, i64, vp]),
    "dbs_ising_init_packed_strided": (i32, [vp, i32, i64, u64, i64, i64, vp]),
    "dbs_ising_sweeps_strided": (i32, [vp, i32, i64, i32, vp, i32, vp, u64, i64, i64, i64, vp, i32, vp]),
    "dbs_ising_pack": (i32, [vp, vp, i32, i64, vp]),
    "dbs_ising_unpack": (i32, [vp, vp, i32, i64, vp]),
    "dbs_ising_sweeps": (i32, [vp, i32, i64, i32, vp, i32, vp, u64, i64, i64, vp, i32, vp]),
    "dbs_ising_site_update": (i32, [vp, i32, i64, i32, i32, vp, C.c_uint32, vp]),
    "dbs_ising_observables": (i32, [vp, i32, i64, f32, f32, vp, vp, vp]),
    "dbs_ising_observables_f32": (i32, [vp, i32, i64, f32, f32, vp, vp, vp]),
    "dbs_ising_correlation": (i32, [vp, i32, i64, vp, vp, vp, vp]),
    "dbs_ising_measure": (i32, [vp, i32, i64, f32, f32, vp, vp, i32, vp, vp]),
    "dbs_ising_set_correlation_method": (i32, [i32]),
    "dbs_observable_accumulate": (i32, [vp, vp, i32, i64, vp, vp, vp]),
    "dbs_em_step": (i32, [vp, vp, vp, i64, i64, f32, f32, f32, f32, i32, u64, i64, i64, vp]),
    "dbs_gaussian_fill": (i32, [vp, i64, i64, u64, i64, i64, vp]),
    "dbs_lincomb": (i32, [vp, i64, i32, vp, vp, i32, vp]),
    "dbs_lincomb_dev": (i32, [vp, i64, i32, vp, vp, i32, vp]),
    "dbs_step_table_next": (i32, [vp, vp, vp, i32, vp]),
    "dbs_em_step_dev": (i32, [vp, vp, i64, i64, vp, vp, vp]),
    "dbs_corrector_coef": (i32, [vp, i64, i64, C.c_double, vp, vp, vp]),
    "dbs_corrector_update": (i32, [vp, vp, vp, i64, i64, vp, vp, vp, vp]),
    "dbs_lj_energy_forces": (i32, [vp, i64, i32, f32, f32, f32, f32, f32, vp, vp, vp]),
    "dbs_mala_propose": (i32, [vp, vp, vp, vp, i64, i32, f32, f32, f32, f32, u64, i64, i64, vp]),
    "dbs_mala_accept": (i32, [vp, vp, vp, vp, vp, vp, vp, i64, i32, f32, f32, u64, i64, i64, vp, vp]),
    "dbs_baoab_first": (i32, [vp, vp, vp, vp, i64, i32, f32, f32, f32, f32, f32, u64, i64, i64, vp]),
    "dbs_baoab_second": (i32, [vp, vp, i64, i32, f32, f32, vp]),
    "dbs_rdf_histogram": (i32, [vp, i64, i32, f32, vp, i32, vp, vp]),
    "dbs_unet_plan_create": (i32, [C.POINTER(UnetConfig), C.POINTER(Param), i32, vp, C.POINTER(vp)]),
    "dbs_unet_plan_destroy": (None, [vp]),
    "dbs_unet_workspace_bytes": (i64, [vp, i64, i32, i32, i32]),
    "dbs_unet_forward": (i32, [vp, vp, vp, i64, vp, i64, i32, i32, i32, vp, i64, vp]),
    "dbs_unet_forward_tap": (i32, [vp, vp, vp, i64, vp, i64, i32, i32, i32, vp, i64, C.c_char_p, vp, i64, vp]),
    "dbs_csn_plan_create": (i32, [C.POINTER(CsnConfig), C.POINTER(Param), i32, vp, vp, C.POINTER(vp)]),
    "dbs_csn_plan_destroy": (None, [vp]),
    "dbs_csn_workspace_bytes": (i64, [vp, i64, i32, i64]),
    "dbs_csn_forward": (i32, [vp, vp, vp, i64, vp, i64, i32, i32, vp, i64, vp]),
    "dbs_conv_tc": (i32, [vp, i32, vp, i32, vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, vp]),
    "dbs_conv_block_tc": (i32, [vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, i64, i32, i32, i32, vp]),
}

for _name, (_res, _args) in PROTOTYPES.items():
    _fn = getattr(lib, _name)  # AttributeError here = library and header out of sync
    _fn.restype = _res
    _fn.argtypes = _args


class EngineError(RuntimeError):
    """An ABI call returned a non-zero status; the text is dbs_last_error()."""


def last_error():
    msg = lib.dbs_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc, what=""):
    if rc != 0:
        kind = {ERR_INVALID: ValueError, ERR_UNSUPPORTED: NotImplementedError}.get(rc, EngineError)
        raise kind(f"{what or 'dbs_b200'}: {last_error()} (status {rc})")


def launch_count():
    return int(lib.dbs_launch_count())


def profile(on):
    check(lib.dbs_profile_enable(1 if on else 0))


def profile_read(reset=True):
    """{class name: (milliseconds, launches)} accumulated since the last reset."""
    n = lib.dbs_profile_num_classes()
    ms = (C.c_double * n)()
    cnt = (C.c_uint64 * n)()
    check(lib.dbs_profile_read(ms, cnt, n, 1 if reset else 0), "profile_read")
    return {lib.dbs_profile_class_name(i).decode(): (float(ms[i]), int(cnt[i])) for i in range(n)}
