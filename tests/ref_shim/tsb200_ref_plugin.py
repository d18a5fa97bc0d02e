"""pytest plugin for the reference-suite run (loaded with `-p tsb200_ref_plugin`):
  * tensors created without an explicit device land on cuda:0 (several reference tests build CPU tensors
    unconditionally; the package under test is GPU-only),
  * every test starts from the same RNG seed (the reference's tests draw unseeded randn inputs),
  * tests listed in tests/ref_suite_xfail.json (currently none) are marked xfail with the recorded reason,
  * the outcome of every test id is written to $TSB200_REF_REPORT as JSON."""
import fnmatch
import json
import os
from pathlib import Path

import pytest
import torch

_XFAIL = json.loads((Path(__file__).resolve().parent.parent / "ref_suite_xfail.json").read_text())
_OUT = {}


def pytest_configure(config):
    torch.set_default_device("cuda:0")


def pytest_collection_modifyitems(config, items):
    for item in items:
        nid = item.nodeid.split("/")[-1]
        for pat, reason in _XFAIL.items():
            if fnmatch.fnmatch(nid, pat):
                item.add_marker(pytest.mark.xfail(reason=reason, strict=False))
                break


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(20240923)
    yield


def pytest_runtest_logreport(report):
    if report.when == "call" or (report.when == "setup" and report.outcome != "passed"):
        nid = report.nodeid.split("/")[-1]
        outcome = report.outcome
        if hasattr(report, "wasxfail"):
            outcome = "xfailed" if report.outcome == "skipped" else "xpassed"
        _OUT[nid] = outcome


def pytest_sessionfinish(session, exitstatus):
    path = os.environ.get("TSB200_REF_REPORT")
    if path:
        Path(path).write_text(json.dumps(_OUT, indent=1, sort_keys=True))
