"""Build the reference's own C++ CPU voxel ops into oracle/_ref/ -- TEST INFRASTRUCTURE ONLY.

Sources are compiled where they lie under /root/reference/projects/BEVFusion/bevfusion/ops/voxel/src
(voxelization.cpp, voxelization_cpu.cpp, scatter_points_cpu.cpp; no WITH_CUDA).  Nothing is copied into git:
oracle/_ref/ is git-ignored (it still travels to the GPU box with the snapshot).

Two modules are produced:
  ref_voxel_asis   the three files exactly as shipped.  Its hard_voxelize_cpu allocates the lookup table
                   {gz,gy,gx} but indexes [x][y][z] (voxelization_cpu.cpp:129-130 vs :75,83), so it is only
                   memory-safe on cubic grids; tests use it on cubic grids only.
  ref_voxel_fixed  the same sources with that ONE allocation changed to {gx,gy,gz} (applied with a text
                   substitution into a temporary directory outside the repo, so no reference source ever lands in the tree) so the reference
                   algorithm can run on the 1440x1440x41 nuScenes grid; this is the timed CPU voxelize arm.

Since round 2 the reference's CUDA sources are built as well (for sm_100a, nvcc cross-compiles here without a GPU):
  ref_bev_pool_cuda  bev_pool.cpp + bev_pool_cuda.cu, unmodified: bev_pool_forward / bev_pool_backward (K1 / K2)
  ref_voxel_cuda     voxelization.cpp + voxelization_cuda.cu + scatter_points_cuda.cu (+ the two CPU files the
                     dispatcher links), unmodified, -DWITH_CUDA: hard_voxelize (deterministic and not),
                     dynamic_voxelize, dynamic_point_to_voxel_forward / _backward on the GPU
They are the GPU-side oracle of the `-m gpu` tests (partial -1 rows of dynamic_voxelize, set equality of the
non-deterministic voxelizer, dynamic scatter) and the same-box "kernel to beat" bench.py times next to ours.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.path.join(os.environ.get("BEVFRONT_REFERENCE_ROOT", "/root/reference"),
                       "projects", "BEVFusion", "bevfusion", "ops", "voxel", "src")
OUT = os.path.join(HERE, "_ref")

_ASIS = "-at::ones({grid_size[2], grid_size[1], grid_size[0]}, coors.options());"
_FIXED = "-at::ones({grid_size[0], grid_size[1], grid_size[2]}, coors.options());"


def built(name):
    d = os.path.join(OUT, name)
    return os.path.isdir(d) and any(f.startswith(name) and f.endswith(".so") for f in os.listdir(d))


def build(verbose=False):
    if not os.path.isdir(REF_SRC):
        print(f"[oracle/_ref] {REF_SRC} not present; using prebuilt files if any", file=sys.stderr)
        return False
    from torch.utils.cpp_extension import load

    os.makedirs(OUT, exist_ok=True)
    files = ["voxelization.cpp", "voxelization_cpu.cpp", "scatter_points_cpu.cpp"]
    if not built("ref_voxel_asis"):
        bd = os.path.join(OUT, "ref_voxel_asis")
        os.makedirs(bd, exist_ok=True)
        load(name="ref_voxel_asis", sources=[os.path.join(REF_SRC, f) for f in files],
             extra_cflags=["-O2"], build_directory=bd, verbose=verbose)
    if not built("ref_voxel_fixed"):
        import tempfile

        sd = tempfile.mkdtemp(prefix="bevfront_ref_src_")
        with open(os.path.join(REF_SRC, "voxelization_cpu.cpp")) as f:
            text = f.read()
        assert _ASIS in text, "reference source changed; update the substitution"
        with open(os.path.join(sd, "voxelization_cpu_fixed.cpp"), "w") as f:
            f.write(text.replace(_ASIS, _FIXED))
        bd = os.path.join(OUT, "ref_voxel_fixed")
        os.makedirs(bd, exist_ok=True)
        load(name="ref_voxel_fixed",
             sources=[os.path.join(REF_SRC, "voxelization.cpp"), os.path.join(sd, "voxelization_cpu_fixed.cpp"),
                      os.path.join(REF_SRC, "scatter_points_cpu.cpp")],
             extra_include_paths=[REF_SRC], extra_cflags=["-O2"], build_directory=bd, verbose=verbose)
    build_cuda(verbose)
    return True


_NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-D__CUDA_NO_HALF_OPERATORS__",
               "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__", "--expt-relaxed-constexpr"]
POOL_SRC = os.path.join(os.path.dirname(os.path.dirname(REF_SRC)), "bev_pool", "src")


def build_cuda(verbose=False):
    """The reference's CUDA ops for sm_100a (flags of projects/BEVFusion/setup.py:18-29 with the gencode list
    replaced).  ~4 min the first time."""
    if not os.path.isdir(REF_SRC):
        return False
    from torch.utils.cpp_extension import load

    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")   # no GPU here: keep torch from guessing an arch list
    if not built("ref_bev_pool_cuda"):
        bd = os.path.join(OUT, "ref_bev_pool_cuda")
        os.makedirs(bd, exist_ok=True)
        load(name="ref_bev_pool_cuda", sources=[os.path.join(POOL_SRC, f) for f in ("bev_pool.cpp", "bev_pool_cuda.cu")],
             extra_cflags=["-O2"], extra_cuda_cflags=_NVCC_FLAGS, build_directory=bd, verbose=verbose, with_cuda=True)
    if not built("ref_voxel_cuda"):
        bd = os.path.join(OUT, "ref_voxel_cuda")
        os.makedirs(bd, exist_ok=True)
        files = ["voxelization.cpp", "scatter_points_cpu.cpp", "scatter_points_cuda.cu", "voxelization_cpu.cpp",
                 "voxelization_cuda.cu"]
        load(name="ref_voxel_cuda", sources=[os.path.join(REF_SRC, f) for f in files],
             extra_cflags=["-O2", "-DWITH_CUDA"], extra_cuda_cflags=_NVCC_FLAGS + ["-DWITH_CUDA"], build_directory=bd,
             verbose=verbose, with_cuda=True)
    return True


def load_ref(name="ref_voxel_fixed"):
    """Import a prebuilt oracle/_ref module (works on the GPU box: needs only torch + the .so)."""
    import importlib.util

    import torch  # noqa: F401  (the .so links against libtorch)

    d = os.path.join(OUT, name)
    if not os.path.isdir(d):
        return None
    for f in os.listdir(d):
        if f.startswith(name) and f.endswith(".so"):
            spec = importlib.util.spec_from_file_location(name, os.path.join(d, f))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
    return None


if __name__ == "__main__":
    ok = build(verbose="-v" in sys.argv)
    print("oracle/_ref:", "built" if ok else "skipped")
