"""Loading of tests/golden/*.npz (written by oracle/gen_golden.py from the reference itself)."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from oracle import head_oracle as HO
from oracle import ref_harness as RH

GOLDEN = Path(__file__).resolve().parent / "golden"
MM_LOSSES = list(HO.MM_CALL_ORDER)
NLP_LOSSES = list(HO.NLP_CALL_ORDER)


def teacher_head_weight():
    g = torch.Generator().manual_seed(2)
    return RH.head_weight(1) + 0.01 * torch.randn(RH.VOCAB, RH.HIDDEN, generator=g)


class HeadCase:
    def __init__(self, trainer: str, loss_type: str):
        self.trainer, self.loss_type = trainer, loss_type
        z = np.load(GOLDEN / f"head_{trainer}_{loss_type}.npz")
        self.z = z
        self.beta = float(z["beta"])
        self.loss = float(z["loss"])
        self.student_splits = [str(s) for s in z["student_splits"]]
        self.teacher_splits = [str(s) for s in z["teacher_splits"]]
        self.hidden = [torch.from_numpy(z[f"hidden_{i}"]) for i in range(len(self.student_splits))]
        self.dhidden = [torch.from_numpy(z[f"dhidden_{i}"]) for i in range(len(self.student_splits))]
        self.labels = [torch.from_numpy(z[f"labels_{i}"]) for i in range(len(self.student_splits))]
        self.ref_hidden = [torch.from_numpy(z[f"ref_hidden_{i}"]) for i in range(len(self.teacher_splits))]
        self.probe_rows = torch.from_numpy(z["probe_rows"])
        self.dw_rows = torch.from_numpy(z["dw_rows"])
        self.dw_colsum = torch.from_numpy(z["dw_colsum"])
        self.dw_sqnorm = float(z["dw_sqnorm"])
        self.weight = RH.head_weight(1)
        self.ref_weight = teacher_head_weight()

    def oracle_inputs(self, cast=None):
        """student dict / teacher dict in the layout of head_oracle.mm_compute_loss."""
        c = (lambda t: t.to(cast).float()) if cast is not None else (lambda t: t)
        student, rnd = {}, []
        for name, h, y in zip(self.student_splits, self.hidden, self.labels):
            if name.startswith("random"):
                rnd.append(HO.HeadInputs(c(h), y))
            else:
                student[name] = HO.HeadInputs(c(h), y)
        if rnd:
            student["random"] = rnd
        teacher = {name: c(h) for name, h in zip(self.teacher_splits, self.ref_hidden)}
        return student, teacher, c(self.weight), c(self.ref_weight)
