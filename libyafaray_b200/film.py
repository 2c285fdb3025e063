"""Per-GPU films of a tile-sharded render and their sum (BASELINE.json configs[2]).

One process per GPU renders its share of the frame's tiles (integration/include/render/tile_shard_b200.h) into a film of
the full frame size.  A film is the reference's accumulation buffer: per pixel one filter-weight sum and, per layer, one
RGBA sum of weight x colour (ImageFilm::addSample, src/render/imagefilm.cc:771-822).  Because filter splats cross tile
borders, shards are SUMMED, not concatenated -- weights and layers alike -- which is exactly what the reference does
when it merges the ".film" files of several render nodes (ImageFilm::imageFilmLoadAllInFolder,
src/render/imagefilm.cc:1072-1090); the image is layer / weight afterwards (ImageFilm::flush, imagefilm.cc:660-700).

File format = the reference's own (ImageFilm::imageFilmSave / imageFilmLoad, src/render/imagefilm.cc:910-1020,1099-1176):
    "YAF_FILMv4_0_0\\0"  int32 x 10 {computer_node, base_sampling_offset, sampling_offset, width, height, cx0, cx1, cy0,
    cy1, n_layers}  float32 weights[height][width]  n_layers x float32 rgba[height][width][4]
so a summed film written by `write_film` can be loaded back by the reference (film_load_save_mode = load-save).

The sum over processes is ONE collective on one flat buffer (weights and all layers concatenated; 41 MB at 1080p with
one layer): `torch.distributed.reduce(SUM)` -- NCCL over NVLink when the tensors live on the GPUs, gloo in the CPU tests.
PyTorch is plumbing here (device buffer + collective), not the product; there is no kernel of ours on this path.
"""
from __future__ import annotations

import dataclasses
import struct

import numpy as np

MAGIC = b"YAF_FILMv4_0_0\0"
_HEADER = struct.Struct("<10i")


@dataclasses.dataclass
class Film:
    width: int
    height: int
    weights: np.ndarray              # (height, width) float32
    layers: np.ndarray               # (n_layers, height, width, 4) float32, weight x colour sums
    computer_node: int = 0
    base_sampling_offset: int = 0
    sampling_offset: int = 0
    cx0: int = 0
    cy0: int = 0

    def flat(self) -> np.ndarray:
        """weights and layers in one contiguous float32 vector (the buffer that is reduced)."""
        return np.concatenate([self.weights.reshape(-1), self.layers.reshape(-1)]).astype(np.float32, copy=False)

    def with_flat(self, flat: np.ndarray) -> "Film":
        n = self.width * self.height
        flat = np.asarray(flat, dtype=np.float32)
        if flat.size != n + self.layers.size:
            raise ValueError(f"flat film buffer has {flat.size} floats, expected {n + self.layers.size}")
        return dataclasses.replace(self, weights=flat[:n].reshape(self.height, self.width).copy(),
                                   layers=flat[n:].reshape(self.layers.shape).copy())


def film_path(prefix: str, computer_node: int = 0) -> str:
    """ImageFilm::getFilmPath (src/render/imagefilm.cc:900-908)."""
    return f"{prefix} - node {computer_node:04d}.film"


def read_film(path: str) -> Film:
    with open(path, "rb") as f:
        data = f.read()
    if not data.startswith(MAGIC):
        raise ValueError(f"{path}: not a YAF_FILMv4_0_0 film")
    off = len(MAGIC)
    node, base_off, samp_off, w, h, cx0, cx1, cy0, cy1, n_layers = _HEADER.unpack_from(data, off)
    off += _HEADER.size
    if w <= 0 or h <= 0 or n_layers < 0 or cx1 != cx0 + w - 1 or cy1 != cy0 + h - 1:
        raise ValueError(f"{path}: inconsistent film header (w={w} h={h} layers={n_layers} cx={cx0}..{cx1} cy={cy0}..{cy1})")
    need = off + 4 * w * h * (1 + 4 * n_layers)
    if len(data) != need:
        raise ValueError(f"{path}: {len(data)} bytes, expected {need}")
    weights = np.frombuffer(data, dtype="<f4", count=w * h, offset=off).reshape(h, w).copy()
    off += 4 * w * h
    layers = np.frombuffer(data, dtype="<f4", count=n_layers * w * h * 4, offset=off).reshape(n_layers, h, w, 4).copy()
    return Film(w, h, weights, layers, node, base_off, samp_off, cx0, cy0)


def write_film(path: str, film: Film) -> None:
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(_HEADER.pack(film.computer_node, film.base_sampling_offset, film.sampling_offset, film.width, film.height,
                             film.cx0, film.cx0 + film.width - 1, film.cy0, film.cy0 + film.height - 1, film.layers.shape[0]))
        f.write(np.ascontiguousarray(film.weights, dtype="<f4").tobytes())
        f.write(np.ascontiguousarray(film.layers, dtype="<f4").tobytes())


def check_compatible(a: Film, b: Film) -> None:
    """The checks ImageFilm::imageFilmLoad makes before it accepts a film (src/render/imagefilm.cc:932-990)."""
    for name in ("width", "height", "cx0", "cy0"):
        if getattr(a, name) != getattr(b, name):
            raise ValueError(f"films differ in {name}: {getattr(a, name)} vs {getattr(b, name)}")
    if a.layers.shape != b.layers.shape:
        raise ValueError(f"films differ in layers: {a.layers.shape} vs {b.layers.shape}")


def sum_films(films) -> Film:
    """Host-side merge, pixel for pixel what imageFilmLoadAllInFolder does (float adds in list order)."""
    films = list(films)
    out = dataclasses.replace(films[0], weights=films[0].weights.copy(), layers=films[0].layers.copy())
    for other in films[1:]:
        check_compatible(out, other)
        out.weights += other.weights
        out.layers += other.layers
        out.sampling_offset = max(out.sampling_offset, other.sampling_offset)
        out.base_sampling_offset = max(out.base_sampling_offset, other.base_sampling_offset)
    return out


def normalized(film: Film, layer: int = 0) -> np.ndarray:
    """layer / weight where weight > 0 (ImageFilm::flush divides by the weight, imagefilm.cc:672-676); (h, w, 4) float32."""
    w = film.weights[..., None]
    return np.where(w > 0.0, film.layers[layer] / np.where(w > 0.0, w, 1.0), 0.0).astype(np.float32)


def reduce_film(film: Film, dst: int = 0, device=None):
    """Sum the films of all ranks of the default process group onto rank `dst` with ONE collective.

    Every rank passes its own film; returns (Film on rank dst / None elsewhere, seconds spent in the collective incl. the
    copies to and from `device`).  `device` = torch.device('cuda', local_rank) with the NCCL backend, None/cpu with gloo."""
    import time

    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        raise RuntimeError("reduce_film needs an initialised torch.distributed process group")
    rank = dist.get_rank()
    # every rank must bring the same frame: compare the geometry before adding anything up
    geo = torch.tensor([film.width, film.height, film.cx0, film.cy0, film.layers.shape[0]], dtype=torch.int64, device=device)
    lo, hi = geo.clone(), geo.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if not torch.equal(lo, hi):
        raise ValueError(f"rank {rank}: films of the ranks differ in size / origin / layer count: min {lo.tolist()} max {hi.tolist()}")
    use_cuda = device is not None and torch.device(device).type == "cuda"
    if use_cuda:
        torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    buf = torch.from_numpy(film.flat())
    if use_cuda:
        buf = buf.pin_memory().to(device, non_blocking=True)
    dist.reduce(buf, dst=dst, op=dist.ReduceOp.SUM)
    out = None
    if rank == dst:
        out = film.with_flat(buf.cpu().numpy())
    elif use_cuda:
        torch.cuda.synchronize(device)
    return out, time.perf_counter() - t0


def psnr(a: np.ndarray, b: np.ndarray, peak: float = 1.0) -> float:
    """PSNR of two linear images after clipping to [0, peak] (what an 8-bit output would keep)."""
    a = np.clip(np.asarray(a, dtype=np.float64), 0.0, peak)
    b = np.clip(np.asarray(b, dtype=np.float64), 0.0, peak)
    mse = float(np.mean((a - b) ** 2))
    return float("inf") if mse == 0.0 else 10.0 * np.log10(peak * peak / mse)


def write_tga(path: str, film: Film, layer: int = 0, srgb: bool = True) -> None:
    """The normalised layer as an uncompressed 32-bit TGA (bottom-up BGRA, the container the reference's built-in TGA output
    writes).  srgb=True applies the sRGB transfer curve to the colour channels; the reference's own colour-space and badge
    handling (src/image/image_output.cc) is NOT reproduced -- load the summed .film into a stock libYafaRay for that."""
    img = np.clip(normalized(film, layer), 0.0, 1.0).astype(np.float64)
    rgb = img[..., :3]
    if srgb:
        rgb = np.where(rgb <= 0.0031308, 12.92 * rgb, 1.055 * np.power(np.maximum(rgb, 1e-12), 1.0 / 2.4) - 0.055)
    out = np.empty((film.height, film.width, 4), dtype=np.uint8)
    out[..., 0] = np.round(rgb[..., 2] * 255.0)
    out[..., 1] = np.round(rgb[..., 1] * 255.0)
    out[..., 2] = np.round(rgb[..., 0] * 255.0)
    out[..., 3] = np.round(img[..., 3] * 255.0)
    header = struct.pack("<BBBHHBHHHHBB", 0, 0, 2, 0, 0, 0, 0, 0, film.width, film.height, 32, 8)
    with open(path, "wb") as f:
        f.write(header)
        f.write(np.ascontiguousarray(out[::-1]).tobytes())  # film rows run top to bottom, TGA's default origin is bottom left
