"""CPU: libclear_b200.so loads and exports every symbol include/clear_b200.h declares; argument
validation that happens before any launch works without a GPU; the product has no CPU fallback."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
HEADER = (ROOT / "include" / "clear_b200.h").read_text()


def _declared():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(clear_[a-z0-9_]+)\s*\(", body)))


def test_header_symbols_are_exported_and_bound():
    from multimodal_unlearning_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/clear_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names, "ctypes binding list and header disagree"
    assert lib.clear_abi_version() == int(re.search(r"#define CLEAR_ABI_VERSION (\d+)", HEADER).group(1))


def test_struct_layouts_match_header_field_order():
    from multimodal_unlearning_b200 import _lib
    for cname, cls in (("ClearGemmArgs", _lib.GemmArgs), ("ClearProjectArgs", _lib.ProjectArgs),
                       ("ClearRowLossArgs", _lib.RowLossArgs), ("ClearBackwardArgs", _lib.BackwardArgs),
                       ("ClearAdamWArgs", _lib.AdamWArgs), ("ClearHeadArgs", _lib.HeadArgs)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), HEADER, flags=re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(",")
            first = names[0].split()[-1].lstrip("*")
            fields.append(first)
            fields += [n.strip().lstrip("*") for n in names[1:]]
        fields = [re.sub(r"\[\d+\]$", "", f) for f in fields]   # `void* dweight_peers[8]` -> dweight_peers
        assert [f[0] for f in cls._fields_] == fields, cname


def test_argument_errors_are_reported_before_any_launch():
    from multimodal_unlearning_b200 import _lib
    lib = _lib.load()
    assert lib.clear_head_vtiles(32064) == 126 and lib.clear_head_vtiles(32000) == 125 and lib.clear_head_vtiles(1) == 1
    assert lib.clear_saliency_chunk_elems() > 0
    assert lib.clear_head_project(None, None) == -1
    a = _lib.ProjectArgs()
    assert lib.clear_head_project(C.byref(a), None) == -1
    assert b"null" in lib.clear_last_error().lower()
    r = _lib.RowLossArgs(logits=16, lse_part=16, loss_rows=16, n_rows=4, vocab=10, logits_ld=16, kind=99)
    assert lib.clear_head_rowloss(C.byref(r), None) == -4
    r = _lib.RowLossArgs(logits=16, lse_part=16, loss_rows=16, n_rows=4, vocab=10, logits_ld=16, kind=_lib.KIND_NPO)
    assert lib.clear_head_rowloss(C.byref(r), None) == -1  # teacher missing
    r = _lib.RowLossArgs(logits=16, lse_part=16, loss_rows=16, n_rows=4, vocab=100, logits_ld=16, kind=_lib.KIND_ENTROPY)
    assert lib.clear_head_rowloss(C.byref(r), None) == -2  # ld < vocab
    g = _lib.GemmArgs(a=16, b=16, d=16, lda=8, ldb=8, ldd=8, m=8, n=8, k=8, cta_group=3)
    assert lib.clear_gemm_bf16(C.byref(g), None) == -4
    g = _lib.GemmArgs(a=16, b=16, d=16, lda=7, ldb=8, ldd=8, m=8, n=8, k=7, cta_group=1)
    assert lib.clear_gemm_bf16(C.byref(g), None) == -3
    assert lib.clear_saliency_accumulate(8, 16, 8, None) == -3
    w = _lib.AdamWArgs(param=16, grad=16, exp_avg=16, exp_avg_sq=16, n=8, step=0)
    assert lib.clear_masked_adamw(C.byref(w), None) == -2


def test_no_cpu_fallback():
    from multimodal_unlearning_b200 import head, salun
    h = torch.zeros(8, 16, dtype=torch.bfloat16)
    w = torch.zeros(32, 16, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="CUDA"):
        head.head_loss_fwd_bwd(h, w, head.KIND_ENTROPY)
    with pytest.raises(RuntimeError, match="CUDA"):
        salun.SaliencyState(salun.FlatLayout([("a", (8,))]), "cpu")


def test_product_never_imports_the_oracle():
    pkg = ROOT / "multimodal_unlearning_b200"
    for f in pkg.glob("*.py"):
        src = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_dweight_slabs_partition_the_vocabulary():
    """Owner slabs of the fused reduce-scatter: contiguous, whole 256-row output tiles, cover [0, V) exactly."""
    import ctypes as C
    from multimodal_unlearning_b200 import _lib
    lib = _lib.load()
    for vocab in (32064, 32000, 13, 256, 257):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for rank in range(world):
                r0, r1 = C.c_int32(), C.c_int32()
                _lib.check(lib.clear_head_dweight_slab(vocab, world, rank, 0, C.byref(r0), C.byref(r1)), "slab")
                assert r0.value == min(prev, vocab) and r0.value <= r1.value <= vocab
                assert r0.value % 256 == 0 or r0.value == vocab
                prev = r1.value
            assert prev == vocab
    assert lib.clear_head_dweight_slab(100, 9, 0, 0, C.byref(C.c_int32()), C.byref(C.c_int32())) < 0


def test_plain_c_consumer_links_and_runs(tmp_path):
    """tests/c/abi_consumer.c: the header compiles as C11 (-Wall -Wextra -Werror), the library links without Python or torch,
    and argument errors come back as negative ClearStatus codes before any launch."""
    import shutil
    import subprocess
    from multimodal_unlearning_b200 import build as B
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    lib = B.build()
    exe = tmp_path / "abi_consumer"
    cmd = [gcc, "-std=c11", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "c" / "abi_consumer.c"),
           "-o", str(exe), f"-L{lib.parent}", f"-l:{lib.name}", f"-Wl,-rpath,{lib.parent}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    assert run.returncode == 0 and run.stdout.strip().endswith("ok"), run.stdout + run.stderr
