"""Issue-rate probes of the SM's FP32 / INT32 pipes (fnb200MeasureAluPeak): the roofline denominator (mode 0) and the
instruction-mix probes behind the kernel design decisions. Prints lane-operations per second and, for the mixes, the SM
scheduler cycles one repetition takes (so that "does an FFMA2 cost one issue slot or two" can be read off directly).

    python tools/alu_probe.py            # on the GPU box
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastnoise2_b200 as fn  # noqa: E402

MODES = [
    (0, "FFMA", 8, "8 FFMA"),
    (1, "IMAD", 8, "8 IMAD"),
    (2, "LOP3 (2 per op)", 8, "16 LOP3"),
    (3, "FFMA + 2 LOP3", 16, "8 FFMA + 16 LOP3"),
    (4, "FFMA2", 8, "4 FFMA2"),
    (5, "4 FFMA2 + 8 LOP3", 12, "4 FFMA2 + 8 LOP3"),
    (6, "4 FFMA2 + 16 LOP3", 16, "4 FFMA2 + 16 LOP3"),
    (7, "4 FFMA2 + 4 LOP3", 12, "4 FFMA2 + 4 LOP3: 8 cycles if an FFMA2 is one issue slot, 12 if two"),
    (8, "4 FFMA2(imm) + 4 LOP3", 12, "the same with immediate multiplier / addend"),
    (9, "8 FFMA + 4 LOP3", 12, "scalar counterpart: 12 issue slots"),
    (10, "4 FFMA2(3 pairs) + 4 LOP3", 12, "FFMA2 with three register-pair operands (6 register reads)"),
]


def main():
    if not fn.device_available():
        raise SystemExit("no CUDA device")
    peak = fn.lib.fnb200MeasureAluPeak(0, 4000)
    for mode, name, per_rep, what in MODES:
        rate = fn.lib.fnb200MeasureAluPeak(mode, 4000)
        # one repetition per thread = per_rep lane-ops; a scheduler issues 32 lanes per cycle at the FFMA rate
        cycles = per_rep * peak / rate if rate else float("nan")
        print("%-24s %7.2f T lane-ops/s   %5.2f scheduler cycles per repetition (%s)" % (name, rate / 1e12, cycles, what))


if __name__ == "__main__":
    main()
