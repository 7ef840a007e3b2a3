"""world_size=2 checks of the sharding rule and the single collective, on CPU with the gloo backend."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      SYNFERENCE_B200_QUIET="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from synference_b200 import distributed as D
    from synference_b200.utils import write_container
    assert D.rank_world() == (rank, world)
    a, b = D.shard_bounds(total, rank, world)
    full = torch.arange(total * 3, dtype=torch.float32).reshape(total, 3)
    local = full[a:b].clone()                      # what this rank "synthesised"
    got = D.gather_rows(local, total)
    assert torch.equal(got, full), f"rank {rank}: gathered tensor differs"
    # per-rank library shards (columns = galaxies), merged on the host like combine_rank_files
    write_container(os.path.join(tmp, f"lib_{rank}.hdf5"),
                    {"Grid/Photometry": full[a:b].numpy().T.copy(), "Grid/Parameters": full[a:b, :2].numpy().T.copy()},
                    {"FilterCodes": ["a", "b", "c"], "ParameterNames": ["redshift", "log_mass"], "rank": rank,
                     "world_size": world})
    D.barrier()
    if rank == 0:
        out = D.merge_rank_shards([os.path.join(tmp, f"lib_{r}.hdf5") for r in range(world)], os.path.join(tmp, "lib.hdf5"))
        from synference_b200.utils import load_library_from_hdf5
        lib = load_library_from_hdf5(out)
        assert np.array_equal(lib["photometry"], full.numpy().T)
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 7])
def test_shard_gather_merge_world2(tmp_path, total):
    mp.spawn(_worker, args=(2, _free_port(), total, str(tmp_path)), nprocs=2, join=True)
