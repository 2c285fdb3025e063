"""Per-GPU films and their sum: the reference's .film format, the merge arithmetic of
ImageFilm::imageFilmLoadAllInFolder (src/render/imagefilm.cc:1072-1090), the one-collective reduce on a world_size-2
gloo group, and -- where the patched reference is prebuilt -- real tile-sharded renders whose summed films are compared
with the film of one process rendering every tile."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from libyafaray_b200 import film

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "integration", "_build")
RENDER_BENCH = os.path.join(BUILD, "render_bench")
needs_render_bench = pytest.mark.skipif(not os.path.exists(RENDER_BENCH), reason="integration/_build/render_bench not prebuilt")


def make_film(seed, w=37, h=23, layers=2, mask=None):
    rng = np.random.default_rng(seed)
    weights = rng.random((h, w), dtype=np.float32) * 3.0
    data = rng.random((layers, h, w, 4), dtype=np.float32)
    if mask is not None:
        weights = weights * mask
        data = data * mask[None, :, :, None]
    return film.Film(w, h, weights.astype(np.float32), data.astype(np.float32), 0, 0, 4, 0, 0)


def test_film_file_round_trip_and_header_layout(tmp_path):
    f = make_film(1)
    p = str(tmp_path / "a - node 0000.film")
    assert film.film_path(str(tmp_path / "a")) == p
    film.write_film(p, f)
    raw = open(p, "rb").read()
    # imagefilm.cc:1113-1123: magic string + NUL, ten ints, then the weights
    assert raw[:15] == b"YAF_FILMv4_0_0\0"
    assert np.frombuffer(raw, "<i4", 10, 15).tolist() == [0, 0, 4, 37, 23, 0, 36, 0, 22, 2]
    assert len(raw) == 15 + 40 + 4 * 37 * 23 * (1 + 4 * 2)
    g = film.read_film(p)
    assert np.array_equal(g.weights, f.weights) and np.array_equal(g.layers, f.layers)
    assert (g.width, g.height, g.sampling_offset) == (37, 23, 4)


def test_film_reader_rejects_damaged_files(tmp_path):
    f = make_film(2)
    p = str(tmp_path / "x.film")
    film.write_film(p, f)
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-4])
    with pytest.raises(ValueError):
        film.read_film(p)
    open(p, "wb").write(b"NOT_A_FILM" + raw[10:])
    with pytest.raises(ValueError):
        film.read_film(p)
    with pytest.raises(ValueError):
        film.sum_films([f, make_film(3, w=38)])


def test_sum_and_normalise_follow_the_reference_merge():
    mask = np.zeros((23, 37), np.float32)
    mask[:, :20] = 1.0
    a, b = make_film(4, mask=mask), make_film(5, mask=1.0 - mask)
    b.sampling_offset = 9
    s = film.sum_films([a, b])
    assert np.array_equal(s.weights, a.weights + b.weights) and np.array_equal(s.layers, a.layers + b.layers)
    assert s.sampling_offset == 9                          # imagefilm.cc:1092: the larger offset is kept
    img = film.normalized(s, 1)
    assert np.array_equal(img[:, :20], a.layers[1][:, :20] / a.weights[:, :20, None])
    empty = film.Film(4, 3, np.zeros((3, 4), np.float32), np.ones((1, 3, 4, 4), np.float32))
    assert np.array_equal(film.normalized(empty), np.zeros((3, 4, 4), np.float32))  # no weight, no colour (and no NaN)
    assert film.psnr(img, img) == float("inf")


def test_tga_writer_round_trips_through_the_test_suite_reader(tmp_path):
    from tests.test_render import read_tga
    f = film.Film(5, 3, np.ones((3, 5), np.float32), np.zeros((1, 3, 5, 4), np.float32))
    f.layers[0, 0, 0] = [1.0, 0.0, 0.0, 1.0]     # top-left pixel red
    f.layers[0, 2, 4] = [0.0, 0.0, 0.5, 1.0]     # bottom-right pixel half blue (linear)
    p = str(tmp_path / "x.tga")
    film.write_tga(p, f)
    img = read_tga(p)                             # rows as stored: bottom-up, BGRA
    assert img.shape == (3, 5, 4)
    assert img[2, 0].tolist() == [0.0, 0.0, 255.0, 255.0]
    assert img[0, 4, 0] == 188.0 and img[0, 4, 3] == 255.0   # sRGB(0.5) = 0.7354 -> 188


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _reduce_worker(rank, world, port, out_dir, mismatch):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = make_film(10 + rank, w=38 if (mismatch and rank == 1) else 37)
    try:
        total, seconds = film.reduce_film(mine, dst=0, device=None)
        if rank == 0:
            film.write_film(os.path.join(out_dir, "total.film"), total)
        else:
            assert total is None
        assert seconds >= 0.0
    except ValueError as e:
        open(os.path.join(out_dir, f"error{rank}.txt"), "w").write(str(e))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_film_reduce_equals_host_sum(tmp_path):
    world = 2
    mp.spawn(_reduce_worker, args=(world, _free_port(), str(tmp_path), False), nprocs=world, join=True)
    got = film.read_film(str(tmp_path / "total.film"))
    want = film.sum_films([make_film(10), make_film(11)])
    assert np.array_equal(got.weights, want.weights) and np.array_equal(got.layers, want.layers)


def test_two_rank_gloo_film_reduce_refuses_different_frames(tmp_path):
    world = 2
    mp.spawn(_reduce_worker, args=(world, _free_port(), str(tmp_path), True), nprocs=world, join=True)
    for rank in range(world):
        assert "differ" in open(str(tmp_path / f"error{rank}.txt")).read()
    assert not os.path.exists(str(tmp_path / "total.film"))


def _render_film(tmp_path, name, integrator, shard, threads=2, aa=2, size=(160, 100), extra=(), accel="yafaray-kdtree-original"):
    prefix = str(tmp_path / name)
    cmd = [RENDER_BENCH, accel, integrator, "48", str(size[0]), str(size[1]), str(aa), prefix + ".tga", str(threads),
           f"tile_shard={shard}", "film_save=" + prefix, *extra]
    p = subprocess.run(cmd, cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, errors="replace", timeout=600)
    assert p.returncode == 0 and "RENDER_BENCH" in p.stdout, p.stdout[-2000:]
    if accel == "b200-kdtree":
        assert "no usable accelerator" not in p.stdout, "libb200rt failed to build the scene on this box"
        assert "wavefront rays closest=" in p.stdout and "per-ray calls outside fibers: 0" in p.stdout, "the render did not go through the wavefront ray queue"
    return film.read_film(film.film_path(prefix))


@needs_render_bench
@pytest.mark.parametrize("count", [2, 3, 8])
def test_tile_shards_of_the_reference_renderer_sum_to_the_whole_frame(tmp_path, count):
    """Stock CPU kd-tree, direct lighting (no order-dependent sampler state in this scene): every pixel is rendered by
    exactly one shard, shards only overlap in the filter border, and the summed film equals the unsharded one up to the
    float rounding of adding the same splats in a different order."""
    whole = _render_film(tmp_path, "whole", "directlighting", "0/1")
    shards = [_render_film(tmp_path, f"s{i}", "directlighting", f"{i}/{count}") for i in range(count)]
    covered = sum((s.weights > 0).astype(np.int32) for s in shards)
    assert covered.min() >= 1, "a pixel was rendered by no shard"
    share = [float((s.weights > 0).mean()) for s in shards]
    assert max(share) < 1.0 / count + 0.35, f"shards are not a partition of the tiles: {share}"
    total = film.sum_films(shards)
    assert np.allclose(total.weights, whole.weights, rtol=1e-5, atol=1e-6)
    for layer in range(whole.layers.shape[0]):
        assert film.psnr(film.normalized(total, layer), film.normalized(whole, layer)) > 100.0


@needs_render_bench
def test_render_sharded_tool_two_processes_gloo(tmp_path):
    """tools/render_sharded.py end to end on CPU: two torchrun ranks, stock kd-tree, gloo film reduce, PSNR against one process."""
    import json
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "render_sharded.py"), "--accelerator", "yafaray-kdtree-original",
           "--integrator", "directlighting", "--cells", "48", "--width", "160", "--height", "100", "--aa", "2", "--threads", "2", "--compare",
           "--save", str(tmp_path / "sum.film")]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, errors="replace", timeout=900)
    lines = [l for l in p.stdout.splitlines() if l.startswith('{"tool": "render_sharded"')]
    assert p.returncode == 0 and len(lines) == 1, p.stdout[-3000:]
    rec = json.loads(lines[0])
    assert rec["processes"] == 2 and rec["backend"] == "gloo"
    assert rec["pixels_with_weight"] == rec["pixels"] == 160 * 100
    assert rec["psnr_vs_single_db"] > 100.0
    assert film.read_film(str(tmp_path / "sum.film")).width == 160


@pytest.mark.gpu
@needs_render_bench
def test_tile_shards_through_the_b200_accelerator_sum_to_the_whole_frame(tmp_path):
    """The same partition through AcceleratorB200's tile_shard_index / tile_shard_count parameters and the wavefront ray
    queue (one GPU here: the shards run one after the other; tools/render_sharded.py runs them on one GPU each).  Compared
    with the unsharded b200 render AND with the stock CPU kd-tree render of every tile."""
    stock = _render_film(tmp_path, "stock", "directlighting", "0/1")
    whole = _render_film(tmp_path, "whole", "directlighting", "0/1", accel="b200-kdtree")
    shards = [_render_film(tmp_path, f"g{i}", "directlighting", f"{i}/4", accel="b200-kdtree") for i in range(4)]
    covered = sum((s.weights > 0).astype(np.int32) for s in shards)
    assert covered.min() >= 1
    assert max(float((s.weights > 0).mean()) for s in shards) < 0.6
    total = film.sum_films(shards)
    assert np.allclose(total.weights, whole.weights, rtol=1e-5, atol=1e-6)
    assert np.allclose(total.weights, stock.weights, rtol=1e-5, atol=1e-6)
    # renderTile() per wavefront block draws the area-light jitter from another RNG stream than renderTile() per tile
    # (tests/test_render.py): the stated bar is 50 dB; measured values are printed
    a, b = film.psnr(film.normalized(total), film.normalized(whole)), film.psnr(film.normalized(total), film.normalized(stock))
    print(f"sharded b200 vs unsharded b200: {a:.1f} dB; vs stock kd-tree: {b:.1f} dB")
    assert a > 50.0 and b > 50.0


@pytest.mark.gpu
@needs_render_bench
def test_static_and_nested_instances_render_like_the_reference(tmp_path):
    """static instances (and instances of instances) are uploaded pre-transformed with the reference's
    own matrix code, so the b200 render of a scene with 9 instanced boxes equals the stock kd-tree render."""
    stock = _render_film(tmp_path, "stock", "directlighting", "0/1", size=(240, 150), extra=("instances=9",))
    b200 = _render_film(tmp_path, "b200", "directlighting", "0/1", size=(240, 150), extra=("instances=9",), accel="b200-kdtree")
    plain = _render_film(tmp_path, "plain", "directlighting", "0/1", size=(240, 150))
    assert film.psnr(film.normalized(stock), film.normalized(plain)) < 45.0, "the instances are not visible in this view: the test would prove nothing"
    value = film.psnr(film.normalized(b200), film.normalized(stock))
    print(f"instances: b200 vs stock kd-tree {value:.1f} dB")
    assert value > 50.0


@pytest.mark.gpu
@needs_render_bench
def test_motion_blur_renders_like_the_reference(tmp_path):
    """Bezier motion-blur meshes and moving instances travel through AcceleratorB200 ->
    b200rt_add_mesh_bezier / _moving, and every ray carries Ray::time_ through the wavefront queue (b200rt_job.times): 6 deforming
    boxes + 6 moving pillars.  With the frame time FORCED to one value ("time_forced", integrator_tiled.cc:309) both renders see
    the same geometry for every sample, so the b200 render must equal the stock kd-tree render like the static scenes do; the
    forced times cover before / inside / at the end of the objects' time ranges, and the images at two times must differ.
    With the time drawn per sample the two renders agree to the Monte Carlo noise of 4 samples per pixel (the sample times come
    from a per-tile RNG, which the block-wise evaluation of the wavefront queue seeds differently)."""
    frames = {}
    for label, t in (("t0", 0.0), ("t37", 0.37), ("t100", 1.0)):
        extra = ("motion=6", "b:time_forced=1", f"f:time_forced_value={t}")
        stock = _render_film(tmp_path, "stock" + label, "directlighting", "0/1", size=(240, 150), extra=extra)
        b200 = _render_film(tmp_path, "b200" + label, "directlighting", "0/1", size=(240, 150), extra=extra, accel="b200-kdtree")
        value = film.psnr(film.normalized(b200), film.normalized(stock))
        print(f"motion blur, time forced to {t}: b200 vs stock kd-tree {value:.1f} dB")
        assert value > 50.0
        frames[label] = stock
    assert film.psnr(film.normalized(frames["t0"]), film.normalized(frames["t100"])) < 40.0, "the objects do not move in this view"
    args = dict(size=(240, 150), aa=4, extra=("motion=6",))
    stock = _render_film(tmp_path, "stock", "directlighting", "0/1", **args)
    b200 = _render_film(tmp_path, "b200", "directlighting", "0/1", accel="b200-kdtree", **args)
    value = film.psnr(film.normalized(b200), film.normalized(stock))
    print(f"motion blur, 4 sample times per pixel: b200 vs stock kd-tree {value:.1f} dB")
    assert value > 28.0


@pytest.mark.gpu
@needs_render_bench
def test_sphere_objects_render_like_the_reference(tmp_path):
    """objects of type "sphere" travel through AcceleratorB200 -> b200rt_add_spheres -> the sphere
    branch of the leaf loop; 40 spheres over the field, b200 render against the stock kd-tree render."""
    stock = _render_film(tmp_path, "stock", "directlighting", "0/1", size=(240, 150), extra=("spheres=40",))
    b200 = _render_film(tmp_path, "b200", "directlighting", "0/1", size=(240, 150), extra=("spheres=40",), accel="b200-kdtree")
    plain = _render_film(tmp_path, "plain", "directlighting", "0/1", size=(240, 150))
    assert film.psnr(film.normalized(stock), film.normalized(plain)) < 45.0, "the spheres are not visible in this view"
    value = film.psnr(film.normalized(b200), film.normalized(stock))
    print(f"spheres: b200 vs stock kd-tree {value:.1f} dB")
    assert value > 50.0


@pytest.mark.gpu
@needs_render_bench
@pytest.mark.parametrize("integrator,extra", [("photonmapping", ("i:diffuse_photons=200000", "i:caustic_photons=20000", "b:finalGather=0")),
                                              ("SPPM", ("i:photons=100000", "i:passNums=2"))])
def test_photon_passes_run_on_fibers(tmp_path, integrator, extra):
    """SURVEY.md 8f row N4 (photon shooting): the reference's own photon worker functions run as logical workers on the fibers
    of the wavefront queue (integration/include/render/photon_fibers_b200.h), so no photon bounce is a one-ray launch; the
    image agrees with the stock render as well as two stock renders agree with each other (photon maps are stochastic:
    radiance sub-sampling and Halton dimensions >= 50 draw from one shared LCG)."""
    import re
    stock = _render_film(tmp_path, "stock", integrator, "0/1", size=(240, 150), aa=1, threads=4, extra=extra)
    again = _render_film(tmp_path, "again", integrator, "0/1", size=(240, 150), aa=1, threads=4, extra=extra)
    prefix = str(tmp_path / "b200")
    cmd = [RENDER_BENCH, "b200-kdtree", integrator, "48", "240", "150", "1", prefix + ".tga", "4", "film_save=" + prefix, *extra]
    p = subprocess.run(cmd, cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, errors="replace", timeout=600)
    assert p.returncode == 0 and "no usable accelerator" not in p.stdout, p.stdout[-2000:]
    stats = re.findall(r"wavefront rays closest=(\d+) .*?per-ray calls outside fibers: (\d+)", p.stdout)
    assert len(stats) >= 2, "expected one statistics line per photon pass and one per render pass"
    assert all(int(per_ray) == 0 for _, per_ray in stats), f"rays were traced one per launch: {stats}"
    assert int(stats[0][0]) > 50000, "the photon pass traced no rays through the queue"
    b200 = film.read_film(film.film_path(prefix))
    floor = film.psnr(film.normalized(again), film.normalized(stock))
    value = film.psnr(film.normalized(b200), film.normalized(stock))
    print(f"{integrator}: b200 vs stock {value:.1f} dB, stock vs stock {floor:.1f} dB")
    assert value > min(floor - 3.0, 40.0)


@pytest.mark.gpu
@needs_render_bench
def test_radiance_pre_gather_runs_on_the_device(tmp_path):
    """SURVEY.md 8f row N4 (photon-map gather): with final gathering on, PhotonIntegrator's radiance-map precompute
    (src/integrator/surface/integrator_photon_mapping.cc:98-147,505-514: one PhotonMap::gather per radiance point) is one batched
    b200pm_gather call (integration/include/photon/photon_gather_b200.h); the gathers are bit-identical to the reference's
    (tests/test_pm.py), so the frame agrees with the stock render as well as two stock renders agree with each other."""
    extra = ("i:diffuse_photons=100000", "i:caustic_photons=10000", "b:finalGather=1", "i:fg_samples=4")
    stock = _render_film(tmp_path, "stock", "photonmapping", "0/1", size=(160, 100), aa=1, threads=4, extra=extra)
    again = _render_film(tmp_path, "again", "photonmapping", "0/1", size=(160, 100), aa=1, threads=4, extra=extra)
    prefix = str(tmp_path / "b200")
    cmd = [RENDER_BENCH, "b200-kdtree", "photonmapping", "48", "160", "100", "1", prefix + ".tga", "4", "film_save=" + prefix, *extra]
    p = subprocess.run(cmd, cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, errors="replace", timeout=600)
    assert p.returncode == 0 and "no usable accelerator" not in p.stdout, p.stdout[-2000:]
    assert "b200pm: radiance pre-gather on the device" in p.stdout, p.stdout[-3000:]
    assert "gathering on the host" not in p.stdout
    b200 = film.read_film(film.film_path(prefix))
    floor = film.psnr(film.normalized(again), film.normalized(stock))
    value = film.psnr(film.normalized(b200), film.normalized(stock))
    print(f"photonmapping with final gather: b200 vs stock {value:.1f} dB, stock vs stock {floor:.1f} dB")
    # six stock renders of this frame agree with each other to 31.1 ... 32.3 dB (build container); a wrong radiance map is far below
    assert value > min(floor - 4.0, 40.0)


@needs_render_bench
@pytest.mark.parametrize("integrator,extra", [("photonmapping", ("i:diffuse_photons=20000", "i:caustic_photons=4000")), ("SPPM", ("i:photons=20000", "i:passNums=2"))])
def test_patched_photon_launch_sites_keep_the_reference_behaviour_on_cpu(tmp_path, integrator, extra):
    """With any accelerator but b200-kdtree, b200::PhotonWorkers::run starts one std::thread per worker -- the reference's own
    code path (integration/include/render/photon_fibers_b200.h): the thread count in the log is the requested one and the
    render completes with a lit image."""
    import re
    prefix = str(tmp_path / "cpu")
    cmd = [RENDER_BENCH, "yafaray-kdtree-original", integrator, "32", "120", "80", "1", prefix + ".tga", "3", "film_save=" + prefix, "i:threads_photons=3", *extra]
    p = subprocess.run(cmd, cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, errors="replace", timeout=600)
    assert p.returncode == 0 and "RENDER_BENCH" in p.stdout, p.stdout[-2000:]
    shots = re.findall(r"Shooting (\d+) photons across (\d+) threads", p.stdout)
    assert shots and all(int(t) == 3 for _, t in shots), shots
    img = film.normalized(film.read_film(film.film_path(prefix)))
    assert float(img[..., :3].mean()) > 0.01 and (film.read_film(film.film_path(prefix)).weights > 0).all()


@pytest.mark.gpu
@needs_render_bench
def test_replaced_material_is_seen_without_a_rebuild(tmp_path):
    """the reference reads visibility LIVE at every hit, the GPU scene bakes it per face, and the scene keeps
    its accelerator when only a material is replaced (src/scene/scene.cc:318).  render_bench renders, replaces the boxes'
    material by an invisible one and renders again: AcceleratorB200::refreshFaceFlags must patch the flags on the device, so the
    second frame has neither boxes nor their shadows -- exactly like the stock kd-tree's second frame."""
    stock = _render_film(tmp_path, "stock", "directlighting", "0/1", size=(240, 150), extra=("rerender_hidden=1",))
    prefix = str(tmp_path / "b200")
    cmd = [RENDER_BENCH, "b200-kdtree", "directlighting", "48", "240", "150", "2", prefix + ".tga", "2", "film_save=" + prefix, "rerender_hidden=1"]
    p = subprocess.run(cmd, cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, errors="replace", timeout=600)
    assert p.returncode == 0 and "no usable accelerator" not in p.stdout, p.stdout[-2000:]
    assert "flags updated on the device" in p.stdout, "the adaptor did not notice the replaced material"
    assert p.stdout.count("AcceleratorB200: Starting build") == 1, "the accelerator was rebuilt: the test no longer exercises the flag update"
    b200 = film.read_film(film.film_path(prefix))
    visible = _render_film(tmp_path, "visible", "directlighting", "0/1", size=(240, 150))
    assert film.psnr(film.normalized(stock), film.normalized(visible)) < 30.0, "hiding the boxes changes nothing in this view"
    value = film.psnr(film.normalized(b200), film.normalized(stock))
    print(f"second frame after the material was replaced: b200 vs stock {value:.1f} dB")
    assert value > 50.0
