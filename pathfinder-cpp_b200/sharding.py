"""Multi-GPU partitioning of the rasterization path (one process per GPU).

The pipeline has no exchange step, so there is no data-path collective inside it. Two partitions exist:

* scenes / frames of a batch: independent -- `scene_share` hands every rank a contiguous share, nothing is exchanged;
* horizontal strips of one large canvas: winding accumulates top-to-bottom per tile column and geometry above the
  view box is never clipped away (core/d3d9/tiler.cpp:143-144, 381-438; shaders/d3d11/bin.comp:120-123), so strip k
  renders the same scene translated by -y0 into a W x (y1 - y0) view box and is exact without its neighbours. Rows are
  contiguous in a row-major RGBA8 framebuffer, so the strips are assembled by ONE all-gather of equal row blocks
  (`gather_strips`), or with no collective at all when every rank's tile kernel stores straight into the presenting
  rank's framebuffer through a peer mapping (`PeerFramebuffer`).

Host logic only: no kernels here, and nothing under oracle/ is touched.
"""
import torch
import torch.distributed as dist

TILE = 16


def strip_bounds(height, world, rank):
    """Rows [y0, y1) of strip `rank`: equal blocks of whole tile rows (the last strips may be shorter or empty)."""
    tile_rows = (int(height) + TILE - 1) // TILE
    per = (tile_rows + world - 1) // world
    y0 = min(rank * per * TILE, int(height))
    y1 = min((rank + 1) * per * TILE, int(height))
    return y0, y1


def strip_rows(height, world):
    """Rows every rank contributes to the all-gather (equal blocks; short strips are padded)."""
    tile_rows = (int(height) + TILE - 1) // TILE
    return ((tile_rows + world - 1) // world) * TILE


def scene_share(n_scenes, world, rank):
    """Contiguous share of a batch of independent scenes: range(begin, end)."""
    per, extra = divmod(int(n_scenes), world)
    begin = rank * per + min(rank, extra)
    return range(begin, begin + per + (1 if rank < extra else 0))


def gather_strips(local, full=None, group=None):
    """All-gather of equal row blocks. local: (rows, width, 4) uint8 on this rank; returns (world * rows, width, 4).
    When `local` already is this rank's slice of `full` the gather is in place (no staging copy)."""
    world = dist.get_world_size(group)
    if full is None:
        full = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(full.view(-1), local.contiguous().view(-1), group=group)
    return full


class PeerFramebuffer:
    """A framebuffer on the presenting rank that every rank can store into (NVLink peer mapping through torch's
    symmetric memory). Rank r's tile kernel is pointed at `strip_ptr(r)`, so the strip gather happens inside the
    kernel's own 16-byte stores; `barrier()` is the only synchronisation."""

    def __init__(self, height, width, world, rank, device, group=None, root=0):
        import torch.distributed._symmetric_memory as symm_mem

        self.rows = strip_rows(height, world)
        self.width, self.world, self.rank, self.root = int(width), world, rank, root
        nbytes = self.rows * world * self.width * 4
        self.local = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.handle = symm_mem.rendezvous(self.local, group if group is not None else dist.group.WORLD)
        self.root_ptr = int(self.handle.buffer_ptrs[root])

    def strip_ptr(self, rank=None):
        rank = self.rank if rank is None else rank
        return self.root_ptr + rank * self.rows * self.width * 4

    def pitch(self):
        return self.width * 4

    def barrier(self):
        self.handle.barrier()

    def root_strip(self, rank=None):
        """Rank `rank`'s rows of the presenting rank's framebuffer as a tensor (rows, width, 4) on THIS process: the
        destination of a copy-engine push (`push_strip`)."""
        rank = self.rank if rank is None else rank
        n = self.rows * self.width * 4
        return self.handle.get_buffer(self.root, (n,), torch.uint8, rank * n).view(self.rows, self.width, 4)

    def push_strip(self, local, stream):
        """Device-to-device copy of this rank's finished strip into the presenting rank's framebuffer over NVLink, on
        `stream`. It is a cudaMemcpyAsync between two contiguous buffers: a COPY ENGINE moves the bytes, so the SMs are free
        for the next frame while 7/8 of the canvas funnels through one GPU's NVLink ingress (when the tile kernel stores
        across NVLink itself, its CTAs sit on every SM waiting for the link and nothing else can be scheduled)."""
        with torch.cuda.stream(stream):
            self.root_strip().copy_(local, non_blocking=True)

    def frame(self, height):
        """The assembled frame (valid on the root rank after barrier())."""
        return self.local.view(self.rows * self.world, self.width, 4)[:height]
