"""ctypes mirrors of the C data model (include/hpgmg_level.h, include/hpgmg_mg.h).

The layouts equal the reference's no-MPI build (finite-volume/source/level.h:65-200, mg.h:22-33), so the
same classes describe a level built by libhpgmg_b200.so (vectors are DEVICE pointers) and one built by
the reference itself compiled as a host library by the test suite (vectors are host pointers); ``fluxes`` carries our opaque device handle.  Used by the host layer (api.py) and by the tests to diff block lists entry by entry.
"""
import ctypes as C

STENCIL_MAX_SHAPES = 3


class _IJK(C.Structure):
    _fields_ = [("i", C.c_int), ("j", C.c_int), ("k", C.c_int)]


class _Side(C.Structure):
    _fields_ = [("box", C.c_int), ("i", C.c_int), ("j", C.c_int), ("k", C.c_int),
                ("jStride", C.c_int), ("kStride", C.c_int), ("ptr", C.c_void_p)]


class blockCopy_type(C.Structure):
    _fields_ = [("subtype", C.c_int), ("dim", _IJK), ("read", _Side), ("write", _Side),
                ("_pad", C.c_char * 48)]


assert C.sizeof(blockCopy_type) == 128


class communicator_type(C.Structure):
    _fields_ = [("num_recvs", C.c_int), ("num_sends", C.c_int),
                ("recv_ranks", C.POINTER(C.c_int)), ("send_ranks", C.POINTER(C.c_int)),
                ("recv_sizes", C.POINTER(C.c_int)), ("send_sizes", C.POINTER(C.c_int)),
                ("recv_buffers", C.POINTER(C.c_void_p)), ("send_buffers", C.POINTER(C.c_void_p)),
                ("allocated_blocks", C.c_int * 3), ("num_blocks", C.c_int * 3),
                ("blocks", C.POINTER(blockCopy_type) * 3)]


assert C.sizeof(communicator_type) == 104


class box_type(C.Structure):
    _fields_ = [("global_box_id", C.c_int), ("low", _IJK), ("dim", C.c_int), ("ghosts", C.c_int),
                ("jStride", C.c_int), ("kStride", C.c_int), ("volume", C.c_int), ("numVectors", C.c_int),
                ("vectors", C.POINTER(C.c_void_p)), ("fp_base", C.c_void_p)]


assert C.sizeof(box_type) == 56


class _BC(C.Structure):
    _fields_ = [("type", C.c_int), ("allocated_blocks", C.c_int * STENCIL_MAX_SHAPES),
                ("num_blocks", C.c_int * STENCIL_MAX_SHAPES),
                ("blocks", C.POINTER(blockCopy_type) * STENCIL_MAX_SHAPES)]


_TIMER_NAMES = ["smooth", "apply_op", "residual", "blas1", "blas3", "boundary_conditions",
                "restriction_total", "restriction_pack", "restriction_local", "restriction_unpack",
                "restriction_recv", "restriction_send", "restriction_wait",
                "interpolation_total", "interpolation_pack", "interpolation_local", "interpolation_unpack",
                "interpolation_recv", "interpolation_send", "interpolation_wait",
                "ghostZone_total", "ghostZone_pack", "ghostZone_local", "ghostZone_unpack",
                "ghostZone_recv", "ghostZone_send", "ghostZone_wait", "collectives", "Total"]


class _Timers(C.Structure):
    _fields_ = [(n, C.c_double) for n in _TIMER_NAMES]


class level_type(C.Structure):
    _fields_ = [("h", C.c_double), ("active", C.c_int), ("num_ranks", C.c_int), ("my_rank", C.c_int),
                ("box_dim", C.c_int), ("box_ghosts", C.c_int),
                ("box_jStride", C.c_int), ("box_kStride", C.c_int), ("box_volume", C.c_int),
                ("numVectors", C.c_int), ("tag", C.c_int), ("boxes_in", _IJK), ("dim", _IJK),
                ("rank_of_box", C.POINTER(C.c_int)), ("num_my_boxes", C.c_int), ("my_boxes", C.POINTER(box_type)),
                ("allocated_blocks", C.c_int), ("num_my_blocks", C.c_int), ("my_blocks", C.POINTER(blockCopy_type)),
                ("boundary_condition", _BC),
                ("exchange_ghosts", communicator_type * STENCIL_MAX_SHAPES),
                ("restriction", communicator_type * 4),
                ("interpolation", communicator_type),
                ("dominant_eigenvalue_of_DinvA", C.c_double), ("must_subtract_mean", C.c_int),
                ("RedBlack_base", C.c_void_p), ("RedBlack_FP", C.c_void_p), ("fluxes", C.c_void_p),
                ("num_threads", C.c_int), ("timers", _Timers),
                ("Krylov_iterations", C.c_int), ("CAKrylov_formations_of_G", C.c_int),
                ("vcycles_from_this_level", C.c_int)]


REFERENCE_SIZEOF_LEVEL = 1296                       # no-MPI reference build, measured
assert C.sizeof(level_type) == REFERENCE_SIZEOF_LEVEL, C.sizeof(level_type)


class _MGTimers(C.Structure):
    _fields_ = [("MGBuild", C.c_double), ("MGSolve", C.c_double)]


class mg_type(C.Structure):
    _fields_ = [("my_rank", C.c_int), ("num_levels", C.c_int), ("levels", C.POINTER(C.POINTER(level_type))),
                ("timers", _MGTimers), ("MGSolves_performed", C.c_int)]


assert C.sizeof(mg_type) == 40


def block_tuple(b):
    """A hashable summary of one blockCopy_type (pointers reduced to 'is a buffer')."""
    return (b.subtype, b.dim.i, b.dim.j, b.dim.k,
            b.read.box, b.read.i, b.read.j, b.read.k, b.read.jStride, b.read.kStride,
            b.write.box, b.write.i, b.write.j, b.write.k, b.write.jStride, b.write.kStride)
