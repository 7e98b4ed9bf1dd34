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
