"""Agent sharding across GPUs: one process per GPU, agents split evenly over the ranks.

Everything before the attention (encoder, policy net, key/query heads) and after it (decoder) is per-agent
independent; the attention needs every agent's key, query and feature map. Each rank writes
those three things for its local agents straight into its own slots of a packed exchange buffer

    [ feature-map region: world x slot_v ]   slot_v(rank)  = bf16 [apr*B][h][w][P*C]
    [ key/query region:   world x slot_kq ]  slot_kq(rank) = keys f32 [apr*B][k_dim] | queries f32 [apr*B][q_dim]

(apr = agents per rank, slots padded to 256 B) and in-place all_gather_into_tensor calls over NCCL / NVLink fill the
other ranks' slots. The exchange is ONE logical step split in two so that it hides: the feature maps (99 % of the bytes)
are gathered asynchronously as soon as the feature encoder has written them and travel while the policy net runs; only
the key/query gather (a few KB per scene) sits on the critical path before the attention. The attention kernel reads
the gathered buffer directly through the per-rank strides of w2c_attn_args; there is no pack or unpack copy.
"""
import torch


def _pad(nbytes, to=256):
    return (nbytes + to - 1) // to * to


class AgentShardLayout:
    def __init__(self, agent_num, world, rank, batch, k_dim, q_dim, h, w, c, planes):
        if agent_num % world:
            raise ValueError("agent_num=%d must be divisible by the number of ranks (%d)" % (agent_num, world))
        if not 0 <= rank < world:
            raise ValueError("rank %d outside world of %d" % (rank, world))
        self.agent_num, self.world, self.rank, self.batch = agent_num, world, rank, batch
        self.apr = agent_num // world
        self.k_dim, self.q_dim = k_dim, q_dim
        self.h, self.w, self.c, self.planes = h, w, c, planes
        rows = self.apr * batch
        self.keys_bytes = rows * k_dim * 4
        self.queries_bytes = rows * q_dim * 4
        self.val_bytes = rows * h * w * planes * c * 2
        # per-rank slots
        self.val_slot_bytes = _pad(self.val_bytes)
        self.queries_off = _pad(self.keys_bytes)                 # inside a key/query slot
        self.kq_slot_bytes = self.queries_off + _pad(self.queries_bytes)
        # regions of the exchange buffer
        self.val_region_off = 0
        self.kq_region_off = world * self.val_slot_bytes
        self.total_bytes = self.kq_region_off + world * self.kq_slot_bytes

    # strides between rank slots, in elements of each sub-array's dtype (what w2c_attn_args wants)
    @property
    def keys_rank_stride(self):
        return self.kq_slot_bytes // 4

    @property
    def queries_rank_stride(self):
        return self.kq_slot_bytes // 4

    @property
    def val_rank_stride(self):
        return self.val_slot_bytes // 2

    @property
    def first_agent(self):
        return self.rank * self.apr

    def allocate(self, device):
        return torch.zeros(self.total_bytes, dtype=torch.uint8, device=device)

    def val_region(self, exchange):
        return exchange[self.val_region_off:self.val_region_off + self.world * self.val_slot_bytes]

    def kq_region(self, exchange):
        return exchange[self.kq_region_off:self.kq_region_off + self.world * self.kq_slot_bytes]

    def views(self, exchange, rank=None):
        """Typed views (keys, queries, val) of one rank's slots of the exchange buffer."""
        r = self.rank if rank is None else rank
        rows = self.apr * self.batch
        kq = self.kq_region(exchange)[r * self.kq_slot_bytes:(r + 1) * self.kq_slot_bytes]
        keys = kq[:self.keys_bytes].view(torch.float32).view(rows, self.k_dim)
        queries = kq[self.queries_off:self.queries_off + self.queries_bytes].view(torch.float32).view(rows, self.q_dim)
        vs = self.val_region(exchange)[r * self.val_slot_bytes:r * self.val_slot_bytes + self.val_bytes]
        val = vs.view(torch.bfloat16).view(rows, self.h, self.w, self.planes * self.c)
        return keys, queries, val

    def dense(self, exchange):
        """Agent-major dense copies (keys, queries, val) of a gathered exchange buffer (tests / debugging)."""
        parts = [self.views(exchange, r) for r in range(self.world)]
        return tuple(torch.cat([p[i] for p in parts], 0) for i in range(3))


def _gather_region(region, slot_bytes, rank, group, async_op=False):
    import torch.distributed as dist
    mine = region[rank * slot_bytes:(rank + 1) * slot_bytes]
    if dist.get_backend(group) == "gloo":
        mine = mine.clone()  # gloo does not document in-place all-gather; the CPU tests take the copy
    return dist.all_gather_into_tensor(region, mine, group=group, async_op=async_op)


def all_gather_values(exchange, layout, group=None, async_op=False):
    """Feature maps of every rank -> every rank. async_op=True returns the work handle: the transfer then runs on
    NCCL's own stream behind whatever the current stream has queued so far, concurrently with what is queued next."""
    return _gather_region(layout.val_region(exchange), layout.val_slot_bytes, layout.rank, group, async_op)


def all_gather_keys_queries(exchange, layout, group=None):
    """Key and query vectors of every rank -> every rank (small; on the critical path before the attention)."""
    return _gather_region(layout.kq_region(exchange), layout.kq_slot_bytes, layout.rank, group)


def all_gather_slots(exchange, layout, group=None):
    """The whole exchange in one go (both regions, synchronously ordered on the current stream)."""
    all_gather_values(exchange, layout, group)
    all_gather_keys_queries(exchange, layout, group)
