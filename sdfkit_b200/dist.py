"""Multi-GPU sharding of the hot path: one process per GPU (torch.distributed for the plumbing).

The grid shards into contiguous z-slabs of CELL LAYERS; a rank samples its own slices plus a halo (one cell
layer below, one above -- recomputed from the analytic SDF, never exchanged), classifies them, and the only
data-path collective is the all-gather of per-slab (vertices, triangles) counts that fixes every rank's global
vertex / triangle offsets (the reference numbers vertices layer-major, so a slab's ids are a contiguous range).
The mesh gather to rank 0 and the row-band image gather are point-to-point copies.  SURVEY.md section 8e.
"""
import ctypes as C

import numpy as np

from . import _native as N
from . import numerics
from .sdf import require_gpu_sdf
from .voxels import GpuMesh, Mesh


def cells_along(n, step):
    """Cells visited along one axis (MarchingCubes.cs:49-68): x = 0, step, ... while x < n - step."""
    return 0 if n <= step else (n - step + step - 1) // step


def partition(n, parts):
    """Contiguous ranges of ceil(n/parts) items, like Vec3Data.PartitionVertically (VectorData.cs:512-526)."""
    size = (n + parts - 1) // parts if parts > 0 else n
    out, lo = [], 0
    for _ in range(parts):
        hi = min(n, lo + size)
        out.append((lo, max(lo, hi)))
        lo = max(lo, hi)
    return out


def weighted_partition(weights, parts):
    """Contiguous ranges of `weights` (one entry per cell layer) with near-equal total weight: cut k goes where the
    running sum crosses k/parts of the total.  Every part gets at least one layer while layers remain."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    if parts <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * max(0, parts - 1)
    if n <= parts:
        return [(k, k + 1) for k in range(n)] + [(n, n)] * (parts - n)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for k in range(1, parts):
        target = cum[-1] * k / parts
        c = int(np.searchsorted(cum, target, side="left"))
        if c > 0 and abs(cum[c - 1] - target) <= abs(cum[min(c, n)] - target):
            c -= 1
        c = max(c, cuts[-1] + 1)                  # at least one layer per part ...
        c = min(c, n - (parts - k))               # ... and leave one for each remaining part
        cuts.append(max(c, cuts[-1]))
    cuts.append(n)
    return [(cuts[k], cuts[k + 1]) for k in range(parts)]


# cost of one active cell (compact + emit) in units of one voxel (sample + classify), measured on B200 at 1024^3:
# (0.17 + 0.90 ms) / 3.9e6 cells  vs  (2.63 + 0.77 + 0.17 ms) / 1.07e9 voxels
ACTIVE_CELL_COST = 82.0
# the same when every rank also delivers its share of the mesh to host memory (e2e): an active cell then costs its compact +
# emit time plus ~60 bytes (1 vertex, 2 triangles) over PCIe at ~54 GB/s = 1.3 ns, a voxel of the distance-only path 1.1 ps
ACTIVE_CELL_COST_E2E = 1200.0


def plan_layers(sdf, vmin, vmax, nx, ny, nz, parts, step=1, clip=True, coarse=128, active_cell_cost=ACTIVE_CELL_COST):
    """Cut the cell layers into `parts` contiguous slabs of near-equal COST instead of equal thickness.  A z-slab job is
    as slow as its busiest rank, and the surface of a scene is rarely spread evenly in z (the README scene fills 18 % of
    the layers).  The per-layer work is estimated from a coarse meshing pass of the same SDF on this GPU (every rank
    computes the same plan; nothing is exchanged): cost(layer) = voxels(layer) + ACTIVE_CELL_COST * active_cells(layer).
    One implementation for every host: sdfk_plan_layers (csrc/sdfk_multi.inl), which the multi-GPU context also uses."""
    sdf = require_gpu_sdf(sdf)
    vmin, vmax = numerics.vec3(vmin), numerics.vec3(vmax)
    out = (C.c_int * (2 * int(parts)))()
    N.check(N.lib().sdfk_plan_layers(sdf.ctx.handle, sdf.handle, N.fptr(vmin), N.fptr(vmax), int(nx), int(ny), int(nz), int(step),
                                     1 if clip else 0, int(parts), float(active_cell_cost), out))
    return [(out[2 * k], out[2 * k + 1]) for k in range(int(parts))]


def slab_slices(kb, ke, step, nz):
    """Voxel slices [z_begin, z_end) a rank must hold to mesh cell layers [kb, ke): its own planes plus the ghost
    layer below (owner of the vertices on the shared plane) and above (contributes to their normals)."""
    ncz = cells_along(nz, step)
    if ke <= kb:
        return 0, 1
    k0 = max(kb - 1, 0)
    k1 = min(ke + 1, ncz)
    return k0 * step, k1 * step + 1


def exclusive_offsets(counts):
    """counts: int64[world, 2] (vertices, triangles per rank) -> (int64[world, 2] exclusive prefix, totals[2])."""
    counts = np.asarray(counts, dtype=np.int64).reshape(-1, 2)
    excl = np.zeros_like(counts)
    excl[1:] = np.cumsum(counts, axis=0)[:-1]
    return excl, counts.sum(axis=0)


class SlabMesher:
    """One rank's share of Sdf.ToMesh on an nx*ny*nz grid: cell layers [kb, ke)."""

    def __init__(self, sdf, vmin, vmax, nx, ny, nz, kb, ke, clip=True, iso=0.0, step=1, colors=True):
        self.colors = bool(colors)      # False: distance-only voxels, vertex colours evaluated from the SDF (Sdf.ToMesh)
        self.sdf = require_gpu_sdf(sdf)
        self.ctx = self.sdf.ctx
        self.min, self.max = numerics.vec3(vmin), numerics.vec3(vmax)
        self.dims = (int(nx), int(ny), int(nz))
        self.kb, self.ke, self.clip, self.iso, self.step = int(kb), int(ke), bool(clip), float(iso), int(step)
        self.z0, self.z1 = slab_slices(self.kb, self.ke, self.step, self.dims[2])
        self.vox = None
        self.mesh = None
        M, Nn = numerics.mesh_transforms(self.min, self.max, *self.dims)
        self.M, self.Nn = N.f32c(M), N.f32c(Nn)

    def sample(self):
        """K1 on this rank's slices (allocates on first use, then re-samples in place)."""
        L = N.lib()
        if self.vox is None:
            h = C.c_void_p()
            nx, ny, nz = self.dims
            fn = L.sdfk_voxels_sample_slab if self.colors else L.sdfk_voxels_sample_distances
            N.check(fn(self.ctx.handle, self.sdf.handle, N.fptr(self.min), N.fptr(self.max), nx, ny, nz,
                       1 if self.clip else 0, self.z0, self.z1, C.byref(h)))
            self.vox = h
        else:
            N.check(L.sdfk_voxels_resample(self.vox, self.sdf.handle, 1 if self.clip else 0))

    def classify(self):
        """K2-K4a; returns this rank's (vertices, triangles)."""
        if self.mesh is not None:
            self.mesh.destroy()
        h = C.c_void_p()
        nv, nt = C.c_int64(), C.c_int64()
        N.check(N.lib().sdfk_mesh_classify(self.ctx.handle, self.vox, self.iso, self.step, self.kb, self.ke, C.byref(h),
                                           C.byref(nv), C.byref(nt)))
        self.mesh = GpuMesh(h)
        return nv.value, nt.value

    def emit(self, vertex_base, triangle_base):
        """K4b with the global offsets fixed by the count all-gather."""
        N.check(N.lib().sdfk_mesh_emit(self.mesh.handle, int(vertex_base), int(triangle_base), N.fptr(self.M), N.fptr(self.Nn)))
        return self.mesh

    def emit_host(self, vertex_base, triangle_base, chunks=0):
        """K4b in sub-ranges, each finished part streaming to page-locked host memory (sdfk_mesh_emit_host): this rank's share
        of the mesh, with global indices, as a host Mesh (zero-copy view of the handle's buffers)."""
        N.check(N.lib().sdfk_mesh_emit_host(self.mesh.handle, int(vertex_base), int(triangle_base), N.fptr(self.M), N.fptr(self.Nn),
                                            int(chunks)))
        host = self.mesh.host_view()
        self.mesh = None                   # the views own the handle now
        return host

    def voxel_count(self):
        """Voxels this rank owns for throughput accounting (halo slices are not counted)."""
        nx, ny, nz = self.dims
        zlo = self.kb * self.step
        zhi = nz if self.ke >= cells_along(nz, self.step) else self.ke * self.step
        return nx * ny * max(0, zhi - zlo)

    def close(self):
        if self.mesh is not None:
            self.mesh.destroy()
            self.mesh = None
        if self.vox is not None:
            N.lib().sdfk_voxels_destroy(self.vox)
            self.vox = None


def merge_meshes(parts):
    """Concatenate per-slab meshes (already carrying global indices) in rank order -> Mesh."""
    parts = [p for p in parts]
    v = np.concatenate([p.Vertices for p in parts]) if parts else np.zeros((0, 3), np.float32)
    c = np.concatenate([p.Colors for p in parts]) if parts else np.zeros((0, 3), np.float32)
    n = np.concatenate([p.Normals for p in parts]) if parts else np.zeros((0, 3), np.float32)
    t = np.concatenate([p.Triangles for p in parts]) if parts else np.zeros((0,), np.int32)
    nonempty = [p for p in parts if len(p.Vertices)]
    if nonempty:
        mn = np.min(np.stack([p.Min for p in nonempty]), axis=0)
        mx = np.max(np.stack([p.Max for p in nonempty]), axis=0)
    else:
        mn = mx = np.zeros(3, np.float32)
    return Mesh(v, c, n, t, mn.astype(np.float32), mx.astype(np.float32))


def to_mesh_by_slabs(sdf, vmin, vmax, nx, ny, nz, nslabs, clip=True, iso=0.0, step=1):
    """Single-process emulation of an `nslabs`-rank job (the slabs run one after another on this GPU).  Used by
    the tests to prove that the sharded result is identical to the single-GPU one."""
    layers = partition(cells_along(nz, step), nslabs)
    slabs = [SlabMesher(sdf, vmin, vmax, nx, ny, nz, kb, ke, clip, iso, step) for kb, ke in layers]
    counts = []
    for s in slabs:
        s.sample()
        counts.append(s.classify())
    excl, _ = exclusive_offsets(counts)
    parts = []
    for s, (vb, tb) in zip(slabs, excl):
        parts.append(s.emit(vb, tb).download())
        s.close()
    return merge_meshes(parts)


class ShardedMesher:
    """One rank's share of Sdf.ToMesh in an N-rank job.  The cell layers are cut into N * slabs_per_rank contiguous
    slabs dealt round-robin (slab g belongs to rank g % N): with slabs_per_rank > 1 a surface concentrated in a few
    layers (the README scene fills 18 % of z) is spread over all ranks instead of landing on one or two, at the price
    of one extra halo per slab.  Global ids stay layer-major: offsets are exclusive sums over the slabs in order g."""

    def __init__(self, sdf, vmin, vmax, nx, ny, nz, rank, world, slabs_per_rank=1, clip=True, iso=0.0, step=1, balanced=False,
                 colors=True, active_cell_cost=ACTIVE_CELL_COST):
        self.rank, self.world, self.spr = int(rank), int(world), int(slabs_per_rank)
        if balanced:
            layers = plan_layers(sdf, vmin, vmax, nx, ny, nz, self.world * self.spr, step, clip, active_cell_cost=active_cell_cost)
        else:
            layers = partition(cells_along(nz, step), self.world * self.spr)
        self.layers = layers
        self.slab_ids = [self.rank + self.world * s for s in range(self.spr)]
        self.slabs = [SlabMesher(sdf, vmin, vmax, nx, ny, nz, layers[g][0], layers[g][1], clip, iso, step, colors) for g in self.slab_ids]

    def sample_classify(self):
        """K1 + K2..K4a on every local slab; returns int64[slabs_per_rank, 2] (vertices, triangles)."""
        out = np.zeros((self.spr, 2), dtype=np.int64)
        for k, s in enumerate(self.slabs):
            if s.ke > s.kb:
                s.sample()
                out[k] = s.classify()
        return out

    def offsets(self, all_counts):
        """all_counts: int64[world, slabs_per_rank, 2] (all-gathered) -> (this rank's [slabs_per_rank, 2] global
        offsets, totals[2])."""
        c = np.asarray(all_counts, dtype=np.int64).reshape(self.world, self.spr, 2)
        by_slab = c.transpose(1, 0, 2).reshape(self.world * self.spr, 2)      # slab g = rank + world * s
        excl, tot = exclusive_offsets(by_slab)
        return excl[self.slab_ids], tot

    def emit(self, offs):
        for s, (vb, tb) in zip(self.slabs, offs):
            if s.mesh is not None:
                s.emit(vb, tb)

    def emit_host(self, offs, chunks=0):
        """Like emit, but every slab's share of the mesh lands in this rank's page-locked host memory; returns the host Meshes."""
        return [s.emit_host(vb, tb, chunks) for s, (vb, tb) in zip(self.slabs, offs) if s.mesh is not None]

    def voxel_count(self):
        return sum(s.voxel_count() for s in self.slabs)

    def stats(self):
        keys = ("classify_ms", "scan_ms", "compact_ms", "emit_ms")
        tot = dict.fromkeys(keys, 0.0)
        for s in self.slabs:
            if s.mesh is not None:
                st = s.mesh.stats()
                for k in keys:
                    tot[k] += st[k]
        return tot

    def close(self):
        for s in self.slabs:
            s.close()


# ------------------------------------------------------------------------------------------------
# torch.distributed plumbing (backend nccl on GPUs; gloo in the CPU tests)
# ------------------------------------------------------------------------------------------------

def all_gather_int64(values, device=None, group=None):
    """All-gather a small int64 vector from every rank -> int64[world, len(values)] (the data-path collective)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.as_tensor(np.asarray(values, dtype=np.int64).reshape(-1), device=device)
    allc = torch.empty(world * mine.numel(), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allc, mine, group=group)
    return allc.cpu().numpy().reshape(world, -1)


def all_gather_counts(nverts, ntris, device=None, group=None):
    """The one data-path collective: all-gather (vertices, triangles) of every slab -> (excl[world,2], totals[2])."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = torch.tensor([int(nverts), int(ntris)], dtype=torch.int64, device=device)
    allc = torch.empty(world * 2, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allc, mine, group=group)
    return exclusive_offsets(allc.cpu().numpy().reshape(world, 2))


def gather_rows(local, counts, dst=0, group=None):
    """Gather variable-length row blocks (a torch tensor per rank, same trailing shape) to `dst` with point-to-point
    copies at the all-gathered offsets.  counts: rows per rank.  Returns the concatenated tensor on dst, else None."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    counts = [int(c) for c in counts]
    if rank == dst:
        out = torch.empty((sum(counts),) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        offs = np.concatenate([[0], np.cumsum(counts)])
        ops = []
        for r in range(world):
            if counts[r] == 0:
                continue
            view = out[offs[r]:offs[r + 1]]
            if r == dst:
                view.copy_(local)
            else:
                ops.append(dist.P2POp(dist.irecv, view, r, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return out
    if counts[rank] > 0:
        for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, local.contiguous(), dst, group)]):
            w.wait()
    return None


class DeviceArray:
    """Zero-copy view of library-owned device memory for torch (`torch.as_tensor(DeviceArray(...), device='cuda')`)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def mesh_device_tensors(gpu_mesh, device):
    """(vertices, colors, normals [nv,3] float32, triangles [nt,3] int32) aliasing the sdfk_mesh's device buffers."""
    import torch
    nv, nt = gpu_mesh.counts()
    ptrs = [C.c_void_p() for _ in range(4)]
    N.check(N.lib().sdfk_mesh_device_ptrs(gpu_mesh.handle, *[C.byref(p) for p in ptrs]))
    out = []
    for p, (rows, ts) in zip(ptrs, ((nv, "<f4"), (nv, "<f4"), (nv, "<f4"), (nt, "<i4"))):
        if rows == 0 or not p.value:
            out.append(torch.empty((0, 3), dtype=torch.float32 if ts == "<f4" else torch.int32, device=device))
        else:
            out.append(torch.as_tensor(DeviceArray(p.value, (rows, 3), ts), device=device))
    return out
