"""Multi-GPU data parallelism for the hot path: one process per GPU, NCCL over
NVLink 5 / NVSwitch through torch.distributed.

Mirror of the reference's cupyx.distributed surface that the sharded-reduction
config uses: `init_process_group(n_devices, rank, backend='nccl')`
(cupyx/distributed/_init.py:14-91) returning an `NCCLBackend` with
`all_reduce(in_array, out_array, op='sum', stream=None)` etc.
(cupyx/distributed/_nccl_comm.py:60-306).  The reference bootstraps NCCL through
its own TCP store; here torch.distributed's rendezvous (env:// under torchrun, or
tcp://host:port) does that plumbing.
"""
from cupy_b200.distributed._comm import (  # noqa: F401
    NCCLBackend, init_process_group, sharded_sum, sharded_var, combine_moments)
from cupy_b200.distributed import array  # noqa: F401,E402
from cupy_b200.distributed.array import (  # noqa: F401,E402
    DistributedArray, distributed_array, make_2d_index_map, REPLICA, MIN, MAX, SUM, PROD)
