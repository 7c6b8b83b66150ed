"""Stand-in for mpi4py (absent from this image). The reference front end only calls
MPI.COMM_WORLD.Get_rank() to choose a GPU id (PASIf.py:27-30,167). Test infrastructure only."""
import os


class _World:
    def Get_rank(self):
        return int(os.environ.get("LOCAL_RANK", os.environ.get("PASIF_RANK", "0")))

    def Get_size(self):
        return int(os.environ.get("WORLD_SIZE", os.environ.get("PASIF_SIZE", "1")))


class MPI:
    COMM_WORLD = _World()
