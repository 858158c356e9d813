"""Recipe for the reference arm: install the UNMODIFIED reference files the hot path needs into
baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box like the built .so).

    python baseline/install_ref.py          (build container only: needs /root/reference)

The reference is a plain script tree — no setup.py / pyproject.toml, so
`pip install --target baseline/_ref /root/reference` fails with "Neither 'setup.py' nor
'pyproject.toml' found" — and its "install" is the files themselves, copied byte for byte:
    MLM_PLL/main.py      set_dataloader + run_one_epoch: the CPU arm of bench.py --impl reference
    rescore.py           find_best_weight / rescore: the CPU arm of the combiner-only workload (c5)
    util/*.py            ArgParser / parse_config / json_saving imported by both
Two third-party imports of those files are absent from this image and are satisfied by the
stand-ins under baseline/shims/ (ours, committed): `ruamel.yaml` (re-exports PyYAML) and `jiwer`
(cer() on the oracle's C Levenshtein — jiwer's own backend is compiled C++, so a C stand-in is
the fair one).  bench.py reports kind "reference" when baseline/_ref is present and falls back to
the oracle port (kind "port") otherwise.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PLLB_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["MLM_PLL/main.py", "rescore.py", "util/arg_parser.py", "util/config.py", "util/saving.py"]


def install() -> str:
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not found: the reference arm can only be installed in the build container")
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
    with open(os.path.join(DST, "INSTALLED_FROM"), "w") as f:
        f.write(f"{REF}: " + ", ".join(FILES) + " (byte-for-byte copies; see baseline/install_ref.py)\n")
    return DST


if __name__ == "__main__":
    print(install())
    sys.exit(0)
