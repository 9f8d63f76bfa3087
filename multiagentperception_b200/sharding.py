"""Agent sharding across GPUs: one process per GPU, agents split evenly over the ranks, ONE all-gather per forward.

Everything before the attention (encoder, policy net, key/query heads) and after it (decoder) is per-agent
independent; the attention needs every agent's key, query and feature map. Each rank writes
those three things for its local agents straight into its own slot of a packed exchange buffer

    slot(rank) = [ keys  f32 [apr*B][k_dim] | queries f32 [apr*B][q_dim] | feature maps bf16 [apr*B][h][w][P*C] ]

(apr = agents per rank, sub-regions padded to 256 B) and a single in-place all_gather_into_tensor over NCCL /
NVLink fills the other slots. The attention kernel then reads the gathered buffer directly through the per-rank
strides of w2c_attn_args; there is no pack or unpack copy.
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
        self.keys_off = 0
        self.queries_off = _pad(self.keys_bytes)
        self.val_off = self.queries_off + _pad(self.queries_bytes)
        self.slot_bytes = self.val_off + _pad(self.val_bytes)

    # strides between rank slots, in elements of each sub-array's dtype (what w2c_attn_args wants)
    @property
    def keys_rank_stride(self):
        return self.slot_bytes // 4

    @property
    def queries_rank_stride(self):
        return self.slot_bytes // 4

    @property
    def val_rank_stride(self):
        return self.slot_bytes // 2

    @property
    def first_agent(self):
        return self.rank * self.apr

    def allocate(self, device):
        return torch.zeros((self.world, self.slot_bytes), dtype=torch.uint8, device=device)

    def views(self, exchange, rank=None):
        """Typed views (keys, queries, val) of one rank's slot of the exchange buffer."""
        r = self.rank if rank is None else rank
        rows = self.apr * self.batch
        slot = exchange[r]
        keys = slot[self.keys_off:self.keys_off + self.keys_bytes].view(torch.float32).view(rows, self.k_dim)
        queries = slot[self.queries_off:self.queries_off + self.queries_bytes].view(torch.float32).view(rows, self.q_dim)
        val = slot[self.val_off:self.val_off + self.val_bytes].view(torch.bfloat16).view(
            rows, self.h, self.w, self.planes * self.c)
        return keys, queries, val

    def dense(self, exchange):
        """Agent-major dense copies (keys, queries, val) of a gathered exchange buffer (tests / debugging)."""
        parts = [self.views(exchange, r) for r in range(self.world)]
        return tuple(torch.cat([p[i] for p in parts], 0) for i in range(3))


def all_gather_slots(exchange, layout, group=None):
    """The one collective of the forward: every rank contributes its slot, in place."""
    import torch.distributed as dist
    flat = exchange.view(-1)
    mine = exchange[layout.rank]
    if dist.get_backend(group) == "gloo":
        mine = mine.clone()  # gloo does not document in-place all-gather; the CPU tests take the copy
    dist.all_gather_into_tensor(flat, mine, group=group)
