"""`rbp_comm_t`: this process's rank in a one-process-per-GPU job (include/rbp.h).  The library runs every exchange itself;
the host only has to hand the 128-byte id from one rank to the others — here over an already initialised
`torch.distributed` group (any backend), or explicitly with `Comm.create(rank, world, id_bytes)`."""
import ctypes
import os

from . import _ffi


def _locate_nccl():
    """The library dlopens libnccl.so.2 (csrc/comm.cu); point it at the torch-bundled copy unless the caller chose one."""
    if os.environ.get("RBP_NCCL_LIB"):
        return
    try:
        import nvidia.nccl as n
        path = os.path.join(list(n.__path__)[0], "lib", "libnccl.so.2")
        if os.path.exists(path):
            os.environ["RBP_NCCL_LIB"] = path
    except Exception:
        pass


class Comm:
    def __init__(self, handle, rank, world):
        self._h, self.rank, self.world = handle, rank, world
        self._lib = _ffi.lib()

    @staticmethod
    def unique_id():
        _locate_nccl()
        buf = (ctypes.c_uint8 * 128)()
        _ffi.check(_ffi.lib().rbp_comm_unique_id(buf), "rbp_comm_unique_id")
        return bytes(buf)

    @classmethod
    def create(cls, rank, world, uid, device=0):
        assert len(uid) == 128
        _locate_nccl()
        h = ctypes.c_void_p()
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(uid)
        _ffi.check(_ffi.lib().rbp_comm_init(rank, world, buf, device, ctypes.byref(h)), "rbp_comm_init")
        return cls(h, rank, world)

    @classmethod
    def from_torch(cls, dist, device=0):
        """Distribute rank 0's id through `torch.distributed` (object broadcast: works on gloo and nccl groups)."""
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls.create(rank, world, box[0], device)

    def barrier(self):
        _ffi.check(self._lib.rbp_comm_barrier(self._h), "rbp_comm_barrier")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.rbp_comm_destroy(self._h)
            self._h = ctypes.c_void_p()
