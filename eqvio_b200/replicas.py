"""Multi-GPU plumbing of the path: one filter does not shard (a chain of dependent factorisations on one
Sigma), so the parallel axis is *replicas* -- independent sequences / Monte-Carlo VIOSimulator instances.
Instances are split into contiguous blocks, one block per rank (one process per GPU); the data path has no
collective; the per-instance trajectories are brought together by ONE all-gather at the end
(NCCL over NVLink on the GPU box, gloo in the CPU tests).  SURVEY.md section 8(e)."""
import numpy as np

TRAJ_COLS = 11  # stamp, position(3), attitude quaternion wxyz(4), body velocity(3) -- IMUState.csv of the reference


def shard_instances(num_instances: int, world_size: int, rank: int):
    """Contiguous block of instance ids for `rank`; sizes differ by at most one, earlier ranks take the extra."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, extra = divmod(num_instances, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def trajectory_row(stamp, est):
    """One row of a trajectory from a state estimate (eqvio_b200.VIOState)."""
    s = est.sensor
    return np.concatenate([[stamp], s.pose_x, s.pose_q, s.velocity])


def gather_trajectories(local, num_instances: int, group=None, device=None):
    """All-gather of the per-instance trajectories.

    local: dict {instance id: (T, 11) array} for this rank's shard (every instance has the same T).
    Returns an array (num_instances, T, 11) on every rank.  Uses torch.distributed when it is initialised
    (a single all_gather of one padded tensor per rank), otherwise returns the local data (world size 1)."""
    import torch
    import torch.distributed as dist

    ids = sorted(local)
    T = local[ids[0]].shape[0] if ids else 0
    if not (dist.is_available() and dist.is_initialized()):
        out = np.zeros((num_instances, T, TRAJ_COLS))
        for i in ids:
            out[i] = local[i]
        return out
    world = dist.get_world_size(group)
    per_rank = -(-num_instances // world)  # ceil: blocks are padded to a common size for the collective
    Tt = torch.tensor([T], dtype=torch.int64, device=device)
    dist.all_reduce(Tt, op=dist.ReduceOp.MAX, group=group)
    T = int(Tt.item())
    buf = torch.zeros((per_rank, T, TRAJ_COLS + 1), dtype=torch.float64, device=device)
    buf[:, :, TRAJ_COLS] = -1.0  # instance id column, -1 = padding
    for k, i in enumerate(ids):
        buf[k, :, :TRAJ_COLS] = torch.from_numpy(np.ascontiguousarray(local[i])).to(buf.device)
        buf[k, :, TRAJ_COLS] = float(i)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    out = np.zeros((num_instances, T, TRAJ_COLS))
    for p in parts:
        p = p.cpu().numpy()
        for k in range(p.shape[0]):
            i = int(p[k, 0, TRAJ_COLS]) if T else -1
            if i >= 0:
                out[i] = p[k, :, :TRAJ_COLS]
    return out
