"""ctypes loader for libovs_b200.so (the C ABI declared in include/ovs_b200.h).

There is no fallback of any kind: if the library has not been built
(`python -m openvslam_b200.build`, or __graft_entry__.build()) importing a symbol raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "lib", "libovs_b200.so")

OVS_OK = 0
ERR_NAMES = {-1: "INVALID_ARG", -2: "CUDA", -3: "NO_DEVICE", -4: "CAPACITY", -5: "OVERFLOW", -6: "UNSUPPORTED", -7: "NUMERIC"}


class OvsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ovs_b200 error %d (%s): %s" % (code, ERR_NAMES.get(code, "?"), msg))
        self.code = code


class OrbParams(C.Structure):
    _fields_ = [("max_num_keypts", C.c_uint32), ("scale_factor", C.c_float), ("num_levels", C.c_uint32),
                ("ini_fast_thr", C.c_uint32), ("min_fast_thr", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError("libovs_b200.so is not built (%s missing). Run `python -m openvslam_b200.build`; "
                              "there is no CPU fallback." % SO_PATH)
        L = C.CDLL(SO_PATH)
        L.ovs_last_error.restype = C.c_char_p
        L.ovs_version.restype = C.c_char_p
        L.ovs_kernel_launch_count.restype = C.c_uint64
        _lib = L
    return _lib


def check(rc):
    if rc != OVS_OK:
        raise OvsError(rc, lib().ovs_last_error().decode("utf-8", "replace"))


def launch_count():
    return int(lib().ovs_kernel_launch_count())
