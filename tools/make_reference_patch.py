#!/usr/bin/env python
"""Generates integration/roger_cuda_backend.patch: the reference-side change that registers the `cuda` backend
(roger/backend.py BACKENDS) and routes `RogerSetup.step` to roger_b200.integration.cuda_step.  Build container only (needs
the read-only checkout at /root/reference); the patch is committed, `tools/apply_reference_patch.py TARGET` applies it to
a COPY of the reference package (nothing is ever written to /root/reference).

Every edit is anchored on an exact source line and fails loudly if the reference changed."""
import os
import shutil
import subprocess
import sys
import tempfile

REF = "/root/reference/roger"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HOST = '("numpy", "cuda")'
EDITS = {
    "backend.py": [
        ('BACKENDS = ("numpy", "jax")',
         'BACKENDS = ("numpy", "jax", "cuda")\n\n# backends whose host-visible arrays are NumPy arrays (`cuda`: the per-cell step runs on the GPU through roger_b200,\n# `vs.*` stay NumPy mirrors of the device planes)\nHOST_NUMPY_BACKENDS = ("numpy", "cuda")'),
        ('    elif backend_name == "numpy":', '    elif backend_name in HOST_NUMPY_BACKENDS:'),
    ],
    "variables.py": [('    if runtime_settings.backend == "numpy":\n        out.flags.writeable = False',
                      f'    if runtime_settings.backend in {HOST}:\n        out.flags.writeable = False')],
    "core/operators.py": [('if runtime_settings.backend == "numpy":', f'if runtime_settings.backend in {HOST}:')],
    "core/utilities.py": [('if runtime_settings.backend == "numpy":\n    linear_regression', f'if runtime_settings.backend in {HOST}:\n    linear_regression')],
    "core/transport.py": [("    if rs.backend == 'numpy':", f"    if rs.backend in {HOST}:")] * 3,
    "cli/roger_run_base.py": [('type=click.Choice(["numpy", "jax"])', 'type=click.Choice(["numpy", "jax", "cuda"])')],
    "roger.py": [(
        '        with state.timers["diagnostics"]:\n            restart.write_restart(state)\n',
        '        with state.timers["diagnostics"]:\n            restart.write_restart(state)\n\n'
        '        if rs.backend == "cuda":\n'
        '            # the per-cell time step (adaptive time stepping ... storages, numerical error, sanity check) on the GPU;\n'
        '            # hooks, mirrors of vs.*, diagnostics and after_timestep are driven by roger_b200.integration\n'
        '            from roger_b200.integration import cuda_step\n\n'
        '            with state.timers["main"]:\n'
        '                cuda_step(self, state)\n'
        '            if rs.profile_mode:\n'
        '                logger.info(" Time step took {:.2f}s", state.timers["main"].last_time)\n'
        '            return\n',
    )],
}


def apply_edits(pkg):
    for rel, edits in EDITS.items():
        path = os.path.join(pkg, rel)
        txt = open(path).read()
        for old, new in edits:
            if old not in txt:
                raise SystemExit(f"anchor not found in {rel}: {old[:60]!r}")
            txt = txt.replace(old, new, 1)
        open(path, "w").write(txt)


def main():
    out = os.path.join(ROOT, "integration", "roger_cuda_backend.patch")
    with tempfile.TemporaryDirectory() as td:
        for side in ("a", "b"):
            os.makedirs(os.path.join(td, side))
            for rel in EDITS:
                dst = os.path.join(td, side, "roger", rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copy(os.path.join(REF, rel), dst)
        apply_edits(os.path.join(td, "b", "roger"))
        r = subprocess.run(["diff", "-ruN", "a", "b"], cwd=td, capture_output=True, text=True)
        # strip the timestamps diff puts behind the file names (reproducible output)
        lines = []
        for ln in r.stdout.splitlines():
            if ln.startswith(("--- ", "+++ ")):
                ln = ln.split("\t")[0]
            lines.append(ln)
        open(out, "w").write("\n".join(lines) + "\n")
    print(out, sum(1 for ln in lines if ln.startswith(("+", "-")) and not ln.startswith(("+++", "---"))), "changed lines")


if __name__ == "__main__":
    main()
