"""Achieved-error log of the GPU parity tests: every call appends one JSON line to gpurun_out/parity_achieved.jsonl (which
gpurun merges back), scripts/make_parity_md.py turns the log into PARITY.md (the margin under each bound)."""
import json
import os

_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_achieved.jsonl")


def record(family, case, quantity, achieved, bound, note=""):
    try:
        os.makedirs(os.path.dirname(_PATH), exist_ok=True)
        with open(_PATH, "a") as f:
            f.write(json.dumps({"family": family, "case": case, "quantity": quantity, "achieved": float(achieved),
                                "bound": float(bound), "note": note}) + "\n")
    except OSError:
        pass


def log_err(achieved, ref):
    """Automatic log of every max-abs comparison of a parity test, keyed by the running test's id (PYTEST_CURRENT_TEST); the
    reference's max-abs rides along so that PARITY.md can also show the error relative to the output's scale."""
    try:
        import numpy as np
        import torch
        test = os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0]
        r = ref.detach().float().abs().max().item() if torch.is_tensor(ref) else float(np.abs(np.asarray(ref, dtype=np.float64)).max())
        os.makedirs(os.path.dirname(_PATH), exist_ok=True)
        with open(_PATH, "a") as f:
            f.write(json.dumps({"auto": True, "test": test, "achieved": float(achieved), "ref_max_abs": r}) + "\n")
    except Exception:
        pass
