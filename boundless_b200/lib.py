"""ctypes binding of libb200zkp.so (include/b200zkp.h).  Fails loudly: no CPU fallback of any kind."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libb200zkp.so")


class B200Error(RuntimeError):
    pass


class Circuit(C.Structure):
    """b200_circuit (include/b200zkp.h); mirrors the synthetic-circuit header of a seal."""
    _fields_ = [("po2", C.c_uint32), ("w_code", C.c_uint32), ("w_data", C.c_uint32), ("w_accum", C.c_uint32),
                ("kind", C.c_uint32)]

    def __repr__(self):
        return "Circuit(po2=%d, widths=%d/%d/%d, kind=%d)" % (self.po2, self.w_code, self.w_data, self.w_accum, self.kind)


class Task(C.Structure):
    _fields_ = [("task_number", C.c_uint32), ("task_height", C.c_uint32), ("command", C.c_uint32),
                ("n_depends_on", C.c_uint32), ("depends_on", C.c_uint32 * 2),
                ("n_keccak_depends_on", C.c_uint32), ("keccak_depends_on", C.c_uint32 * 2)]


CMD_KECCAK, CMD_FINALIZE, CMD_JOIN, CMD_SEGMENT, CMD_UNION = range(5)

# every symbol include/b200zkp.h declares: (restype, argtypes)
_cp, _vp, _u32, _u64, _sz = C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t
SYMBOLS = {
    "b200_init": (_cp, [C.c_int]),
    "b200_last_error": (_cp, []),
    "b200_device_count": (C.c_int, []),
    "b200_device_pci_bus_id": (_cp, [C.c_int, C.c_char_p, C.c_int]),
    "b200_batch_intt": (_cp, [_vp, _u32, _u32, _vp]),
    "b200_batch_ntt": (_cp, [_vp, _u32, _u32, _vp]),
    "b200_batch_expand_ntt": (_cp, [_vp, _vp, _u32, _u32, _u32, _vp]),
    "b200_batch_zk_shift": (_cp, [_vp, _u32, _u32, _vp]),
    "b200_batch_bit_reverse": (_cp, [_vp, _u32, _u32, _vp]),
    "b200_batch_intt_zk_shift": (_cp, [_vp, _u32, _u32, _vp]),
    "b200_poseidon2_rows": (_cp, [_vp, _vp, _u32, _u32, _vp]),
    "b200_poseidon2_fold": (_cp, [_vp, _vp, _u32, _vp]),
    "b200_merkle_tree": (_cp, [_vp, _vp, _u32, _u32, _vp]),
    "b200_fri_fold": (_cp, [_vp, _vp, _u32, _vp, _vp]),
    "b200_evaluate_scratch_words": (_sz, [_u32, _u32]),
    "b200_batch_evaluate_any": (_cp, [_vp, _vp, _u32, _u32, _vp, _vp, _vp]),
    "b200_shutdown": (_cp, []),
    "b200_mix_poly_coeffs": (_cp, [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _vp]),
    "b200_eltwise_sum_extelem": (_cp, [_vp, _vp, _u32, _u32, _vp]),
    "b200_poly_divide_scratch_words": (_sz, [_u32]),
    "b200_poly_divide": (_cp, [_vp, _u32, _vp, _vp, _vp, _vp]),
    "b200_prefix_products_scratch_words": (_sz, [_u32]),
    "b200_prefix_products": (_cp, [_vp, _u32, _vp, _vp]),
    "b200_eltwise_add_elem": (_cp, [_vp, _vp, _vp, _sz, _vp]),
    "b200_eltwise_copy_elem": (_cp, [_vp, _vp, _sz, _vp]),
    "b200_eltwise_zeroize_elem": (_cp, [_vp, _sz, _vp]),
    "b200_gather_sample": (_cp, [_vp, _vp, _sz, _u32, _sz, _vp]),
    "b200_scatter": (_cp, [_vp, _vp, _u32, _vp, _vp, _u32, _vp]),
    "b200_merkle_open_words": (_sz, [_u32, _u32, _u32]),
    "b200_merkle_open": (_cp, [_vp, _vp, _vp, _u32, _u32, _u32, _u32, _vp]),
    "b200_commit_group": (_cp, [_vp, _vp, _vp, _u32, _u32, _vp]),
    "b200_seal_words": (_sz, [C.POINTER(Circuit)]),
    "b200_prover_create": (_cp, [C.POINTER(_vp), C.c_int, C.POINTER(Circuit), _u32]),
    "b200_prover_destroy": (None, [_vp]),
    "b200_prover_device_bytes": (_sz, [_vp]),
    "b200_prove_segment_async": (_cp, [_vp, _u32, C.POINTER(Circuit), _u64, _vp, _vp]),
    "b200_prefetch_trace_async": (_cp, [_vp, _u32, C.POINTER(Circuit), _vp]),
    "b200_recursion_async": (_cp, [_vp, _u32, C.POINTER(Circuit), _vp, _sz, _vp, _sz, _vp]),
    "b200_recursion_dev_async": (_cp, [_vp, _u32, C.POINTER(Circuit), _vp, _sz, _vp, _sz, _vp]),
    "b200_prove_lift_async": (_cp, [_vp, _u32, C.POINTER(Circuit), _u64, _vp, C.POINTER(Circuit), _vp, _vp, _vp, C.POINTER(C.c_int)]),
    "b200_recursion_verified_async": (_cp, [_vp, _u32, C.POINTER(Circuit), _vp, C.POINTER(Circuit), _vp, C.POINTER(Circuit), _vp, _vp,
                                            C.POINTER(C.c_int)]),
    "b200_seal_to_device": (_cp, [_vp, _u32, _vp, _sz]),
    "b200_prover_query": (C.c_int, [_vp, _u32]),
    "b200_verify_async": (_cp, [_vp, _u32, _vp, _sz, C.POINTER(C.c_int)]),
    "b200_verify_circuit_async": (_cp, [_vp, _u32, C.POINTER(Circuit), _vp, _sz, C.c_int, C.POINTER(C.c_int)]),
    "b200_prover_wait": (_cp, [_vp, _u32]),
    "b200_prover_last_ms": (C.c_float, [_vp, _u32]),
    "b200_prover_mark": (_cp, [_vp, _u32, _u32]),
    "b200_prover_marks_ms": (C.c_float, [_vp, _u32, _u32, _u32, _u32]),
    "b200_witgen_to_host": (_cp, [_vp, _u32, C.POINTER(Circuit), _u64, _vp]),
    "b200_kernel_launches": (_u64, []),
    "b200_host_alloc": (_cp, [C.POINTER(_vp), _sz]),
    "b200_host_free": (None, [_vp]),
    "b200_planner_new": (_vp, []),
    "b200_planner_free": (None, [_vp]),
    "b200_planner_enqueue_segment": (C.c_int64, [_vp]),
    "b200_planner_enqueue_keccak": (C.c_int64, [_vp]),
    "b200_planner_finish": (C.c_int64, [_vp]),
    "b200_planner_task_count": (_sz, [_vp]),
    "b200_planner_get_task": (C.c_int, [_vp, _sz, C.POINTER(Task)]),
    "b200_planner_next_task": (C.c_int, [_vp, C.POINTER(Task)]),
}



class SpparkError(C.Structure):
    """sppark::Error returned BY VALUE by the risc0-sys-compatible exports (include/b200_risc0_sys_compat.h)."""
    _fields_ = [("code", C.c_int32), ("message", C.c_void_p)]      # message: malloc()ed, freed by the caller


# include/b200_risc0_sys_compat.h: the original risc0-sys / sppark symbol names
COMPAT_SYMBOLS = {
    "sppark_init": (SpparkError, []),
    "sppark_batch_iNTT": (SpparkError, [_vp, _u32, _u32]),
    "sppark_batch_NTT": (SpparkError, [_vp, _u32, _u32]),
    "sppark_batch_zk_shift": (SpparkError, [_vp, _u32, _u32]),
    "sppark_batch_expand": (SpparkError, [_vp, _vp, _u32, _u32, _u32]),
    "sppark_poseidon2_rows": (SpparkError, [_vp, _vp, _u32, _u32]),
    "sppark_poseidon2_fold": (SpparkError, [_vp, _vp, _sz]),
    "supra_poly_divide": (SpparkError, [_vp, _sz, _vp, _vp]),
}

_lib = None


def load():
    """Load the in-tree shared library.  Raises B200Error if it has not been built (python -m boundless_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise B200Error("libb200zkp.so not built: run `python -m boundless_b200.build` (needs nvcc; there is no CPU path)")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in list(SYMBOLS.items()) + list(COMPAT_SYMBOLS.items()):
            f = getattr(L, name)           # AttributeError if the ABI and the header drift apart
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(err):
    if err is not None:
        raise B200Error(err.decode() if isinstance(err, bytes) else str(err))


def check_sppark(err):
    """Turn a by-value sppark::Error into an exception, freeing its message the way the Rust Drop does."""
    if err.code == 0 and not err.message:
        return
    msg = C.string_at(err.message).decode() if err.message else "sppark error %d" % err.code
    if err.message:
        C.CDLL(None).free(C.c_void_p(err.message))
    raise B200Error("%s (code %d)" % (msg, err.code))


def require_gpu(device=0):
    L = load()
    n = L.b200_device_count()
    if n <= 0:
        raise B200Error("no CUDA device visible: boundless_b200 has no CPU fallback")
    check(L.b200_init(device))
    return L
