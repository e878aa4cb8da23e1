"""Is the kernel an ncu report captured the kernel the current sources build? (no GPU needed)

    python tools/sass_identity.py gpurun_out/prof_c4_r2l.ncu-rep c4

Rebuilds the workload's cubin through the product's own lowering + NVRTC path and compares its fn2_entry SASS, instruction by
instruction, with the SASS stored in the report (disassembler spelling differences and branch targets normalised away).
"""
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))


def norm(text):
    text = re.sub(r"\s+", " ", text).strip().rstrip(";").strip()
    text = text.replace(".reuse", "").replace("UPT, UPT, ", "").replace("PT, PT, ", "")
    text = re.sub(r"(U?P\d), U?PT, ", r"\1, ", text)  # carry-out predicate pairs: cuobjdump 12.9 spells the unused second one
    text = re.sub(r"0x[0-9a-f]+$", "ADDR", text) if text.split(" ")[0].split(".")[0] in ("BRA", "BSSY", "CALL", "MOV") or " BRA" in text else text
    return text


def main():
    report, workload = sys.argv[1], sys.argv[2]
    import attribute_sass
    text = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text[text.index('"Address"'):])))
    column = rows[0].index("Source")
    captured = [norm(r[column]) for r in rows[1:] if len(r) > column]
    cubin, _, _ = attribute_sass.build_cubin(workload, "strict")
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "fn2_entry", cubin], capture_output=True, text=True).stdout
    current = [norm(m.group(1)) for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(.*?);", sass)]
    for stream in (captured, current):
        while stream and stream[-1].startswith("NOP"):
            stream.pop()
    differing = [(i, a, b) for i, (a, b) in enumerate(zip(current, captured)) if a != b]
    print("%s: %d instructions in the report, %d in the current build, %d differ" % (workload, len(captured), len(current), len(differing) + abs(len(captured) - len(current))))
    for i, a, b in differing[:5]:
        print("   #%d  now: %s   |   captured: %s" % (i, a, b))


if __name__ == "__main__":
    main()
