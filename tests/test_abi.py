"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol the header declares,
the Python mirror keeps the reference's state_dict contract, and the product never routes through the oracle."""
import ctypes
import os
import re

import pytest
import torch

from oracle import stage_oracle as O
from tvqaplus_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "stage_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(stage_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert "stage_cqa_fused_fwd" in names and "stage_linear_fwd" in names
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libstage_b200.so does not export %s" % n
    assert sorted(_lib.SYMBOLS) == names, "ctypes table and header disagree"
    assert _lib.lib().stage_abi_version() == 4


def test_bad_arguments_return_error_codes_not_crashes():
    l = _lib.lib()
    assert l.stage_linear_fwd(None, None) == 1  # STAGE_E_BADARG
    assert b"null" in l.stage_last_error()
    assert l.stage_packed_weight_bytes(500, 128, 64) == 0  # n > 304 unsupported
    assert l.stage_packed_weight_bytes(128, 384, 32) == 12 * 16384
    assert l.stage_packed_weight_bytes(300, 768, 32) == 24 * 2 * 4 * 304 * 16


def test_cpu_tensors_are_rejected_loudly():
    from tvqaplus_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.pack_weight(torch.randn(128, 128))


@pytest.mark.parametrize("heads,n_keys,n_elem", [(0, 72, 680947), (4, 92, 813555)])
def test_state_dict_contract(heads, n_keys, n_elem):
    """Key count / element count of the reference (measured on the live reference)."""
    from tvqaplus_b200.model import STAGE
    m = STAGE(O.default_opt(input_encoder_n_heads=heads, cls_encoder_n_heads=heads))
    sd = m.state_dict()
    assert len(sd) == n_keys and sum(v.numel() for v in sd.values()) == n_elem
    for k in ("input_encoder.stacked_encoderBlocks.0.conv.0.depthwise_conv.weight",
              "input_encoder.stacked_encoderBlocks.0.conv.1.pointwise_conv.weight", "c2q_down_projection.2.weight",
              "classifier.conv.2.weight", "cls_encoder.stacked_encoderBlocks.0.position_encoding.pe",
              "bert_word_encoding_fc.2.weight", "vid_fc.0.weight", "concat_fc.4.bias", "input_embedding.1.weight",
              "cls_projection_layers.0.conv.2.weight", "temporal_scoring_st_layers.0.conv.0.weight"):
        assert k in sd, k
    assert tuple(sd["c2q_down_projection.2.weight"].shape) == (128, 384)
    assert tuple(sd["classifier.conv.2.weight"].shape) == (1, 256)
    assert tuple(sd["input_encoder.stacked_encoderBlocks.0.position_encoding.pe"].shape) == (500, 128)
    # strict load of an oracle-shaped dict round-trips
    m2 = STAGE(O.default_opt(input_encoder_n_heads=heads, cls_encoder_n_heads=heads))
    m2.load_state_dict(sd, strict=True)


def test_position_table_matches_oracle():
    from tvqaplus_b200.model import PositionEncoding
    assert torch.equal(PositionEncoding().pe, O.position_encoding_table())


def test_train_mode_with_dropout_fails_loudly():
    """dropout > 0 in train mode exists only in the training engine (gradients enabled, CUDA tensors): without gradients it
    is refused, and with them a CPU batch fails like every other call - never a silent CPU path."""
    import torch
    from tvqaplus_b200.model import STAGE
    m = STAGE(O.default_opt()).train()
    with torch.no_grad(), pytest.raises(NotImplementedError, match="dropout"):
        m(O.make_synthetic_batch(1, 2, 2, 3, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        m(O.make_synthetic_batch(1, 2, 2, 3, 3))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tvqaplus_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(d, f)).read()
                assert "oracle" not in txt.lower() or f == "stage.py" and "oracle" not in txt, (d, f)


def test_host_helpers():
    from tvqaplus_b200.model import compute_temporal_iou, expand_span, flat_list_of_lists, get_high_iou_sapns
    assert expand_span([2, 5], 3) == [0, 8] and expand_span([7, 9], 2) == [5, 11]
    assert flat_list_of_lists([[1, 2], [3]]) == [1, 2, 3]
    assert compute_temporal_iou([0, 4], [2, 6]) == O.compute_temporal_iou([0, 4], [2, 6]) == 2 / 6
    assert compute_temporal_iou([1, 1], [1, 1]) == 0
    assert get_high_iou_sapns([(0, 4)], [[[0, 3, 0.9], [8, 9, 0.5]]]) == [[(0, 4), [0, 3]]]


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """sizeof / field offsets of every argument struct as gcc sees include/stage_b200.h against the ctypes mirrors."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pairs = [("StageLinearArgs", _lib.LinearArgs), ("StageEncoderArgs", _lib.EncoderArgs), ("StageMhaArgs", _lib.MhaArgs),
             ("StageCqaArgs", _lib.CqaArgs), ("StageHeadArgs", _lib.HeadArgs), ("StageProposalArgs", _lib.ProposalArgs),
             ("StageProbeArgs", _lib.ProbeArgs), ("StageGemmArgs", _lib.GemmArgs)]
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "stage_b200.h"', "int main(void) {"]
    for cname, ct in pairs:
        lines.append('  printf("%s %%zu", sizeof(%s));' % (cname, cname))
        for fname, _ in ct._fields_:
            lines.append('  printf(" %%zu", offsetof(%s, %s));' % (cname, fname))
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    for (cname, ct), line in zip(pairs, out):
        toks = line.split()
        assert toks[0] == cname
        assert int(toks[1]) == ctypes.sizeof(ct), "%s: sizeof %s (C) vs %d (ctypes)" % (cname, toks[1], ctypes.sizeof(ct))
        for (fname, _), off in zip(ct._fields_, toks[2:]):
            assert int(off) == getattr(ct, fname).offset, "%s.%s: offset %s (C) vs %d (ctypes)" % (cname, fname, off, getattr(ct, fname).offset)
