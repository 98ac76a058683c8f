"""CPU-side tests: the C-ABI library builds, loads and exports every symbol
declared in include/ilm_b200.h (no compute calls without a GPU), the product
fails loudly without a device, host logic (grid, containers, column sharding),
the world_size-2 gloo path of the Schur all-gather, and the CPU emulation of
the FFT-convolution kernel bodies."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "immersedlayers.jl_b200", "csrc")


@pytest.fixture(scope="session")
def built_lib():
    so = os.path.join(ROOT, "immersedlayers.jl_b200", "libilm_b200.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-C", CSRC, "-j", str(os.cpu_count() or 4)])
    return so


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "ilm_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ilm_[a-zA-Z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = ctypes.CDLL(built_lib)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    import ilm_b200
    assert declared == set(ilm_b200._lib.SIGNATURES), declared ^ set(ilm_b200._lib.SIGNATURES)


def test_sass_is_sm100a_with_fp64_tensor_and_no_local_fallback(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN3ilm9k_lu_gemmEiiiPKdS1_Pdi", built_lib],
                          capture_output=True, text=True).stdout
    assert "DMMA" in sass          # FP64 tensor pipe in the LU trailing update


def test_no_cpu_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ilm_b200 as ilm
    g = ilm.PhysicalGrid.centered(32)
    with pytest.raises(ilm.IlmError, match="no CUDA device"):
        ilm.SurfaceScalarCache(ilm.bodies.circle(1.0, 1.4 * g.dx), g)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "immersedlayers.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, fn)).read()
                # comments may mention the oracle; importing, including, linking or executing it may not
                bad = re.search(r"import\s+ilm_oracle|from\s+ilm_oracle|#include\s*[\"<][^\">]*oracle|libilm_oracle|"
                                r"sys\.path[^\n]*oracle", src)
                assert not bad, (fn, bad.group(0) if bad else "")


def test_grid_and_containers():
    import ilm_b200 as ilm
    from ilm_b200 import _lib as L
    g = ilm.PhysicalGrid.centered(406, half_width=2.02 * 406 / 406)
    g = ilm.PhysicalGrid(406, 406, 0.01, (203, 203))          # examples/caches.ipynb:62
    assert g.layout_shape(L.NODES_PRIMAL) == (405, 405)
    x, y = g.coordinates(L.NODES_PRIMAL)
    assert abs(x[0] + 2.02) < 1e-12 and abs(x[-1] - 2.02) < 1e-12   # caches.ipynb cell 25
    q = ilm.Edges(g)
    assert len(q) == 406 * 405 * 2 and q.u.shape == (406, 405) and q.v.shape == (405, 406)
    n = ilm.Nodes(ilm.Dual, g)
    n.set(np.arange(406 * 406, dtype=float).reshape((406, 406), order="F"))
    assert n.array()[3, 2] == 3 + 406 * 2
    with pytest.raises(ilm.DimensionMismatch):
        n.set(np.zeros(5))
    v = ilm.VectorData(7)
    assert v.u.shape == (7,) and len(v) == 14


def test_bodies():
    import ilm_b200 as ilm
    x, y, nx, ny, ds = ilm.bodies.ellipse(1.0, 0.5, 0.01)
    assert abs(np.sum(x * nx * ds) - np.pi * 0.5) < 1e-3          # area by the divergence theorem
    assert np.abs(nx ** 2 + ny ** 2 - 1).max() < 1e-12
    x, y, nx, ny, ds = ilm.bodies.rectangle(0.5, 0.25, 0.01)
    assert abs(ds.sum() - 3.0) < 1e-12 and abs(np.sum(x * nx * ds) - 0.5) < 1e-12
    b = ilm.bodies.multibody_c4(4 / 4094)
    assert b[5].shape[0] == 10 and 3900 < b[0].shape[0] < 5400       # BASELINE config C4: ~4000 points


def test_column_ranges():
    from ilm_b200 import shard
    for n in (0, 1, 2, 7, 141, 4593):
        for world in (1, 2, 3, 8):
            r = shard.column_ranges(n, world)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert all(c0 % 2 == 0 for c0, c1 in r if c1 > c0)      # column pairs never straddle ranks
            sizes = [c1 - c0 for c0, c1 in r]
            assert max(sizes) - min(sizes) <= 3                      # pair granularity + odd tail


def _gloo_worker(rank, world, port, n, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from ilm_b200 import shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ranges = shard.column_ranges(n, world)
    full = torch.arange(n * n, dtype=torch.float64)               # column-major N x N reference
    c0, c1 = ranges[rank]
    blk = full[c0 * n:c1 * n].clone()
    got = shard.allgather_columns(blk, n, ranges)

    class FakeCache:
        N = n

    def builder(cache, scale=1.0, cols=None):
        a, b = cols
        return (scale * full[a * n:b * n]).view(b - a, n).t()
    S = shard.create_schur_sharded(builder, FakeCache(), scale=2.0)
    ok = bool(torch.equal(got, full)) and bool(torch.equal(S, 2.0 * full.view(n, n).t()))
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 141])
def test_schur_allgather_gloo_world2(n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n) % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_fft_kernel_bodies_on_cpu():
    """csrc/test_fft_host.cu: the exact __host__ __device__ kernel bodies with 512
    emulated threads per CTA vs long-double DFT / direct convolution."""
    exe = os.path.join(CSRC, "build", "test_fft_host")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-x", "c++", "-I/usr/local/cuda/include", "-pthread",
                           "-o", exe, os.path.join(CSRC, "test_fft_host.cu")])
    # ILM_TEST_BIG: also the radix-2Q bodies for lengths 8192 / 16384 (csrc/ilm_conv_big.cuh)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=900, env=dict(os.environ, ILM_TEST_BIG="1"))
    assert out.returncode == 0 and "PASS" in out.stdout, out.stdout[-2000:]
    assert "Lx=16384" in out.stdout and "Ly=16384" in out.stdout


# ---------------------------------------------------------------- slab decomposition (section 8e, config C5b)
def test_slab_partition_is_consistent():
    """ilm_slab_partition / ilm_slab_counts (pure host arithmetic): the ranks tile the rows and the
    tile columns, row boundaries are even, and what p sends to q is what q expects from p."""
    from ilm_b200 import shard
    for NX, NY in ((24, 20), (24, 600), (600, 24), (130, 129), (4096, 4096)):
        rows = NY - 1
        for world in (1, 2, 3, 8):
            infos = [shard.slab_info(NX, NY, rows, world, r) for r in range(world)]
            assert infos[0].row0 == 0 and infos[-1].row1 == infos[0].MYp >= rows
            assert infos[0].tc0 == 0 and infos[-1].tc1 == infos[0].ntc == infos[0].Lx
            for a, b in zip(infos, infos[1:]):
                assert a.row1 == b.row0 and a.tc1 == b.tc0 and a.whi == b.wlo
            assert all(i.row0 % 2 == 0 and i.row1 % 2 == 0 for i in infos)
            for phase in (0, 1):
                cnt = [shard.slab_counts(NX, NY, i, phase) for i in infos]
                for p in range(world):
                    for q in range(world):
                        assert cnt[p][0][q] == cnt[q][1][p]
                total = sum(sum(c[0]) for c in cnt)
                assert total == 4 * infos[0].ntc * infos[0].MYp          # the whole spectrum, in doubles


def _slab_worker(rank, world, port, NX, NY, q):
    """numpy emulation of the three slab stages around REAL all_to_all_single calls (gloo), with the
    partition, the counts and the block order of csrc/ilm_slab.cu; result vs the oracle."""
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ilm_oracle as o
    from ilm_b200 import lgf, shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mx, my = NX - 1, NY - 1
    info = shard.slab_info(NX, NY, my, world, rank)
    peers = [shard.slab_info(NX, NY, my, world, r) for r in range(world)]
    Lx, Ly, MYp = info.Lx, info.Ly, info.MYp
    G = lgf.lgf_table(max(NX, NY))
    c0, factor = 0.37, 2.5
    w = np.random.default_rng(3).standard_normal((mx, my))
    ref = o.inverse_laplacian(o.ConvPlan(G[:NX, :NY]), w, c0, factor)
    # multiplier on the (2Lx, 2Ly) pad: even extension of (G - c0) / factor
    K = np.zeros((2 * Lx, 2 * Ly))
    ii = np.minimum(np.arange(2 * Lx), 2 * Lx - np.arange(2 * Lx))
    jj = np.minimum(np.arange(2 * Ly), 2 * Ly - np.arange(2 * Ly))
    okx, oky = ii < NX, jj < NY
    K[np.ix_(okx, oky)] = (G[np.ix_(ii[okx], jj[oky])] - c0) / factor
    Ghat = np.fft.fft2(K).real

    def kx_of(tc, ms):                      # tile column -> x frequency (ilm_conv.cuh layout)
        px, a = divmod(tc, Lx // 2)
        return 2 * (2 * a + ms) + px

    # stage 1: FFT_x of my rows, spectrum in tile layout S[tc, row, ms]
    mine = np.zeros((MYp, mx))
    r0, r1 = info.row0, min(info.row1, my)
    mine[r0:r1] = w[:, r0:r1].T
    X = np.fft.fft(mine[info.row0:info.row1], n=2 * Lx, axis=1)          # (my rows, kx)
    S = np.zeros((info.ntc, MYp, 2), dtype=complex)
    for tc in range(info.ntc):
        for ms in (0, 1):
            S[tc, info.row0:info.row1, ms] = X[:, kx_of(tc, ms)]

    def pack(A, blocks):                   # blocks: (tc0, tc1, row0, row1) per peer, peer-major
        out = [A[t0:t1, a:b].reshape(-1) for (t0, t1, a, b) in blocks]
        return np.concatenate(out) if out else np.zeros(0, dtype=complex)

    def exchange(buf, phase):
        send, recv = shard.slab_counts(NX, NY, info, phase)
        sb = torch.from_numpy(buf.view(np.float64).copy())
        assert sb.numel() == sum(send)
        rb = torch.empty(sum(recv), dtype=torch.float64)
        dist.all_to_all_single(rb, sb, output_split_sizes=recv, input_split_sizes=send)
        return rb.numpy().view(complex)

    got = exchange(pack(S, [(p.tc0, p.tc1, info.row0, info.row1) for p in peers]), 0)
    # stage 2: my tile columns, all rows
    C = np.zeros((info.tc1 - info.tc0, MYp, 2), dtype=complex)
    off = 0
    for p in peers:
        n = (info.tc1 - info.tc0) * (p.row1 - p.row0) * 2
        C[:, p.row0:p.row1] = got[off:off + n].reshape(info.tc1 - info.tc0, p.row1 - p.row0, 2)
        off += n
    for t in range(info.tc1 - info.tc0):
        for ms in (0, 1):
            col = np.fft.fft(C[t, :, ms], n=2 * Ly) * Ghat[kx_of(info.tc0 + t, ms)]
            C[t, :, ms] = np.fft.ifft(col)[:MYp]
    got = exchange(pack(C, [(0, info.tc1 - info.tc0, p.row0, p.row1) for p in peers]), 1)
    # stage 3: every tile column, my rows
    S2 = np.zeros((info.ntc, info.row1 - info.row0, 2), dtype=complex)
    off = 0
    for p in peers:
        n = (p.tc1 - p.tc0) * (info.row1 - info.row0) * 2
        S2[p.tc0:p.tc1] = got[off:off + n].reshape(p.tc1 - p.tc0, info.row1 - info.row0, 2)
        off += n
    X2 = np.zeros((info.row1 - info.row0, 2 * Lx), dtype=complex)
    for tc in range(info.ntc):
        for ms in (0, 1):
            X2[:, kx_of(tc, ms)] = S2[tc, :, ms]
    out = np.fft.ifft(X2, axis=1)[:, :mx].real                          # (my rows, mx)
    err = np.abs(out[: r1 - r0] - ref[:, r0:r1].T).max() / np.abs(ref).max() if r1 > r0 else 0.0
    q.put((rank, float(err)))
    dist.destroy_process_group()


@pytest.mark.parametrize("NX,NY", [(24, 20), (24, 600), (300, 24)])
def test_slab_exchange_gloo_world2(NX, NY):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() + NX + NY) % 2000
    procs = [ctx.Process(target=_slab_worker, args=(r, 2, port, NX, NY, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert [r for r, _ in res] == [0, 1]
    assert all(e < 1e-12 for _, e in res), res


def test_julia_shim_binds_only_declared_symbols():
    """julia/ImmersedLayersB200 (the maintainer-side ccall shim of INTEGRATION.md; not executable here:
    no Julia toolchain) may only reference entry points that include/ilm_b200.h declares."""
    import re
    from ilm_b200 import _lib as L
    src = open(os.path.join(ROOT, "julia", "ImmersedLayersB200", "src", "ImmersedLayersB200.jl")).read()
    used = set(re.findall(r"\(:(ilm_[a-z_A-Z0-9]+), lib\)", src)) | set(re.findall(r"_c[23]\(:(ilm_[a-zA-Z_0-9]+)", src))
    assert len(used) >= 35
    header = open(os.path.join(ROOT, "include", "ilm_b200.h")).read()
    for sym in sorted(used):
        assert sym in L.SIGNATURES, sym
        assert re.search(r"\b%s\s*\(" % sym, header), sym
    also = set(re.findall(r"\(:(ilm_[a-z_A-Z0-9]+), lib\)", open(os.path.join(ROOT, "INTEGRATION.md")).read()))
    assert also and all(s in L.SIGNATURES for s in also)


def test_forcing_host_logic():
    """RigidTransform algebra, model constructors and the generated field (src/forcing.jl:14-138, 293-325); no GPU."""
    import ilm_b200 as ilm
    from ilm_b200 import forcing as F
    tr = ilm.RigidTransform((1.0, 2.0), 0.3)
    ident = tr * tr.inv()
    assert abs(ident.x) < 1e-15 and abs(ident.y) < 1e-15 and ident.angle == 0.0
    body = ilm.bodies.rectangle(0.5, 0.25, 0.05)
    moved = ilm.RigidTransform((-1.0, -1.0), np.pi / 4)(body)
    assert np.abs(moved[2] ** 2 + moved[3] ** 2 - 1).max() < 1e-14 and np.array_equal(moved[4], body[4])
    assert abs(np.sum(moved[0] * moved[2] * moved[4]) - 0.5) < 1e-12            # area is invariant
    back = ilm.RigidTransform((-1.0, -1.0), np.pi / 4).inv()(moved)
    assert np.abs(back[0] - body[0]).max() < 1e-14 and np.abs(back[3] - body[3]).max() < 1e-14
    x, y = (ilm.RigidTransform((0.5, 0.0), np.pi / 2) * ilm.RigidTransform((1.0, 0.0), 0.0))(np.array([1.0]), np.array([0.0]))
    assert abs(x[0] - 0.5) < 1e-15 and abs(y[0] - 2.0) < 1e-15                 # b first, then a

    def fcn(*a):
        pass

    m = ilm.AreaForcingModel(fcn)
    assert m.shape is None and m.transform is None
    m = ilm.PointForcingModel((np.zeros(2), np.zeros(2)), fcn, ddftype="m4prime")
    assert isinstance(m.transform, ilm.RigidTransform) and m.kwargs == {"ddftype": "m4prime"}
    with pytest.raises(ilm.MethodError):
        ilm.LineForcingModel(body, tr, 3.0)
    assert ilm.ForcingModelAndRegion(None, None) == []
    g = ilm.PhysicalGrid.centered(128)
    gf = F.GeneratedField(ilm.Nodes(ilm.Primal, g), ilm.SpatialGaussian(0.5, 0.1, 0, 0, 10), g)
    assert abs(gf().numpy().sum() * g.dx ** 2 - 10.0) < 1e-6
    gfe = F.GeneratedField(ilm.Edges(g), [ilm.SpatialGaussian(0.3, 0.3, 0.1, 0, 1), ilm.SpatialGaussian(0.2, 0.4, 0, 0.2, 2)], g)
    assert abs(gfe().u.sum() * g.dx ** 2 - 1.0) < 1e-6 and abs(gfe().v.sum() * g.dx ** 2 - 2.0) < 1e-6


def test_tools_inner_products_and_body_lists():
    """test/tools.jl "Inner Products" (:27-50) and "Operations on body lists" (:55-98) on the host mirror of
    src/tools.jl; the grid integrals are pinned by the values printed in examples/caches.ipynb (cells 32, 34)."""
    import json
    import ilm_b200 as ilm
    from ilm_b200 import tools as T
    g = ilm.PhysicalGrid(406, 406, 0.01, (203, 203))
    w = ilm.Nodes(ilm.Dual, g).fill(1.0)
    u = ilm.Nodes(ilm.Dual, g).fill(1.0)
    assert np.sqrt(T.dot(w, u, g)) == T.norm(w, g)
    area = ((g.NX - 2) * g.dx) ** 2                       # (xlim[2]-xlim[1]) (ylim[2]-ylim[1])
    assert abs(T.norm(w, g) - np.sqrt(area)) < 1e-12 and abs(T.integrate(w, g) - area) < 1e-10
    small = ilm.Nodes(ilm.Dual, ilm.PhysicalGrid(5, 5, 0.01, (2, 2)))
    with pytest.raises(ilm.MethodError):
        T.dot(w, small, g)
    with pytest.raises(ilm.DimensionMismatch):
        T.dot(small, small, g)
    nb = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_notebook_values.json")))
    og = ilm.Nodes(ilm.Primal, g).fill(1.0)
    assert abs(T.integrate(og, g) - nb["caches_integrate_ones_grid"]["values"][0]) < 1e-10
    ovg = T.ones(ilm.Edges(g))
    got = T.integrate(ovg, g)
    assert np.abs(np.array(got) - np.array(nb["caches_integrate_ones_gridgrad"]["values"])).max() < 1e-10
    assert T.integrate(T.ones(ilm.Edges(g), 2), g)[0] == 0.0 and len(T.integrate(T.ones(ilm.EdgeGradient(g)), g)) == 4
    # surface data
    body = ilm.bodies.circle(1.0, 1.4 * g.dx)
    n = len(body[0])
    ds = ilm.ScalarData(n, data=body[4].copy())
    os_ = T.ones(ilm.ScalarData(n))
    assert abs(T.dot(os_, os_, ds) - 2 * np.pi) < 1e-3
    assert abs(T.dot(os_, os_, ds) - nb["caches_dot_ones_surface"]["values"][0]) < 1e-11
    nrm = ilm.VectorData(n, data=np.concatenate([body[2], body[3]]))
    assert abs(T.integrate(T.pointwise_dot(nrm, nrm), ds) - 2 * np.pi) < 1e-3
    # body lists (two copies of the body)
    bl = ilm.bodies.concat(body, body)
    X = ilm.VectorData(n, data=np.concatenate([body[0], body[1]]))
    Xl = ilm.VectorData(2 * n, data=np.concatenate([bl[0], bl[1]]))
    dsl = ilm.ScalarData(2 * n, data=bl[4].copy())
    assert T.dot(Xl, Xl, dsl, bl, 1) == T.dot(Xl, Xl, dsl, bl, 2) == T.dot(X, X, ds)
    assert T.norm(Xl, dsl, bl, 1) == T.norm(Xl, dsl, bl, 2) == T.norm(X, ds)
    nl = ilm.VectorData(2 * n, data=np.concatenate([bl[2], bl[3]]))
    assert abs(T.integrate(T.pointwise_dot(nl, nl), dsl, bl, 2) - 2 * np.pi) < 1e-3
    fl, gl = ilm.ScalarData(2 * n), ilm.ScalarData(2 * n).set(np.random.default_rng(0).random(2 * n))
    T.copyto(fl, gl, bl, 2)
    assert not T.view(fl, bl, 1).any() and np.array_equal(T.view(fl, bl, 2), T.view(gl, bl, 2))
    fl = ilm.ScalarData(2 * n)
    T.copyto(fl, T.view(gl, bl, 2), bl, 2)
    assert not T.view(fl, bl, 1).any() and np.array_equal(T.view(fl, bl, 2), T.view(gl, bl, 2))
    with pytest.raises(ilm.DimensionMismatch):
        T.copyto(fl, [1.0, 2.0, 3.0], bl, 2)
    fv, gv = ilm.VectorData(2 * n), ilm.VectorData(2 * n).set(np.random.default_rng(1).random(4 * n))
    T.copyto(fv, gv, bl, 1)
    assert not T.view(fv, bl, 2)[0].any() and not T.view(fv, bl, 2)[1].any()
    assert np.array_equal(T.view(fv, bl, 1)[0], T.view(gv, bl, 1)[0]) and np.array_equal(T.view(fv, bl, 1)[1], T.view(gv, bl, 1)[1])
    # the added-mass integral of examples/neumann.ipynb cell 37 has this form: -integrate(df o nrm, ds, bl, 2)
    assert isinstance(T.integrate(nl, dsl, bl, 2), list)


def test_tools_cache_forms():
    """dot / norm / integrate / view with a cache argument (src/cache.jl:815-909): GridScaling weighs by dx^2 / ds,
    IndexScaling by plain sums, integrate always uses the areas.  The cache is a stand-in without a device plan."""
    import ilm_b200 as ilm
    from ilm_b200 import tools as T

    class HostOnlyCache(ilm.SurfaceScalarCache):
        def __init__(self, body, g, scaling):
            self.g, self.scaling, self.N, self.device = g, scaling, len(body[0]), False
            self.areas_ = np.asarray(body[4])
            self.body_first = np.asarray(body[5])

        def __del__(self):
            pass

    g = ilm.PhysicalGrid(64, 64, 0.05, (32, 32))
    body = ilm.bodies.circle(1.0, 0.07)
    bl = ilm.bodies.concat(body, body)
    n = len(body[0])
    cg, ci = HostOnlyCache(bl, g, ilm.GridScaling), HostOnlyCache(bl, g, ilm.IndexScaling)
    w = ilm.Nodes(ilm.Primal, g).fill(2.0)
    assert T.dot(w, w, cg) == T.dot(w, w, g) and abs(T.dot(w, w, ci) - T.dot(w, w, g) / g.dx ** 2) < 1e-9
    assert T.integrate(w, cg) == T.integrate(w, ci) == T.integrate(w, g)
    f = ilm.ScalarData(2 * n).fill(1.0)
    ds = ilm.ScalarData(2 * n, data=bl[4].copy())
    assert T.dot(f, f, cg) == T.dot(f, f, ds) and T.dot(f, f, ci) == 2.0 * n
    assert T.dot(f, f, cg, 2) == T.dot(f, f, ds, bl, 2) and T.dot(f, f, ci, 2) == float(n)
    assert T.integrate(f, ci, 1) == T.integrate(f, ds, bl, 1) and T.norm(f, ci, 1) == np.sqrt(n)
    assert T.view(f, cg, 2).shape == (n,)


def test_operators_reject_containers_of_another_grid():
    """every operator checks that its grid / point containers belong to the cache's grid before
    the C call (the library copies ilm_layout_size doubles through the raw pointer).  No GPU needed: the check
    fires before the plan is touched."""
    import ilm_b200 as ilm
    api = ilm.api

    class FakeCache(api.SurfaceScalarCache):
        def __init__(self, g, N):                      # no plan: only what the checks read
            self.g, self.N = g, N

        def __del__(self):
            pass

    g, gs = ilm.PhysicalGrid.centered(32), ilm.PhysicalGrid.centered(16)
    cache = FakeCache(g, 10)
    P, D, E = ilm.Nodes(ilm.Primal, gs), ilm.Nodes(ilm.Dual, gs), ilm.Edges(gs)
    Pg, Dg, Eg = ilm.Nodes(ilm.Primal, g), ilm.Nodes(ilm.Dual, g), ilm.Edges(g)
    f, fbad = ilm.ScalarData(10), ilm.ScalarData(7)
    calls = [
        (api.divergence, (P, Eg, cache)), (api.divergence, (Pg, E, cache)), (api.grad, (E, Pg, cache)),
        (api.curl, (E, Dg, cache)), (api.curl, (D, Eg, cache)), (api.laplacian, (D, D, cache)),
        (api.inverse_laplacian, (E, cache)), (api.surface_divergence, (P, f, cache)), (api.surface_grad, (f, P, cache)),
        (api.surface_curl, (D, f, cache)), (api.surface_curl, (fbad, Dg, cache)), (api.regularize, (Eg, ilm.VectorData(9), cache)),
        (api.interpolate, (ilm.VectorData(10), E, cache)), (api.regularize_normal, (E, f, cache)),
        (api.normal_interpolate, (fbad, Eg, cache)), (api.convective_derivative, (E, Eg, cache)),
        (api.w_cross_v, (Eg, D, Eg, cache)), (api.mask, (ilm.XEdges(gs), cache)), (api.complementary_mask, (P, cache)),
        (api.regularize_normal, (ilm.EdgeGradient(gs), ilm.VectorData(10), cache)),
        (api.normal_interpolate, (ilm.VectorData(10), ilm.EdgeGradient(gs), cache)),
        (api.regularize_normal_dot, (Eg, ilm.TensorData(9), cache)),
    ]
    for fn, args in calls:
        with pytest.raises(ilm.DimensionMismatch):
            fn(*args)


def test_column_range_rule_matches_python_sharding():
    """ilm_column_range (the partition the library uses for ilm_create_schur_sharded / ilm_dirichlet_poisson) is the
    rule of shard.column_ranges: contiguous, even-sized blocks that cover [0, n) (no GPU needed)."""
    import ctypes as C
    import ilm_b200 as ilm
    from ilm_b200 import shard
    lib = ilm._lib.load()
    for n in (0, 1, 2, 7, 141, 4593, 4594):
        for world in (1, 2, 3, 4, 8):
            ref = shard.column_ranges(n, world)
            got = []
            for r in range(world):
                lo, hi = C.c_int(), C.c_int()
                assert lib.ilm_column_range(n, world, r, C.byref(lo), C.byref(hi)) == 0
                got.append((lo.value, hi.value))
            assert got == ref
            assert got[0][0] == 0 and got[-1][1] == n and all(a[1] == b[0] for a, b in zip(got, got[1:]))
            assert all(lo % 2 == 0 for lo, _ in got if lo < n)
    lo, hi = C.c_int(), C.c_int()
    assert lib.ilm_column_range(10, 2, 2, C.byref(lo), C.byref(hi)) != 0


def test_comm_entry_points_fail_cleanly_without_a_plan():
    import ilm_b200 as ilm
    lib = ilm._lib.load()
    assert lib.ilm_comm_init(None, None, 0, 0, 1) == ilm._lib.EINVAL
    assert lib.ilm_comm_destroy(None) == ilm._lib.EINVAL


def test_compact_intfact_table_grows_with_the_fourier_number():
    """The direct-table form of the IF-HERK stage complements cuts the plan_intfact table where it has decayed to
    rounding level; a fixed 64-entry cut truncated it silently for Fourier numbers above ~30 (ADVICE, round 1)."""
    from ilm_b200 import timemarching as tm
    small = tm._compact_intfact_table(0.5, 2048)
    large = tm._compact_intfact_table(50.0, 2048)
    assert small.shape[0] == 64 and large.shape[0] > 64
    for T in (small, large):
        assert abs(T[-1, 0]) <= 1e-17 * abs(T[0, 0])
    # never larger than the grid: nothing is cut then
    assert tm._compact_intfact_table(50.0, 80).shape[0] == 80
