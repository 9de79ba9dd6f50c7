"""Markdown table of one bench.py JSON line (headline + secondary list).   python tools/bench_table.py <file.json>"""
import json
import sys


def fmt(v):
    if v is None:
        return "-"
    return "%.1f M" % (v / 1e6) if v >= 1e6 else "%.0f k" % (v / 1e3)


def main():
    d = json.loads([ln for ln in open(sys.argv[1]) if ln.startswith("{")][-1])
    print("| configuration | value (device-resident input) | e2e (pinned host input) | CPU arm | beam kernel | streaming stage | transcripts identical |")
    print("|---|---|---|---|---|---|---|")
    cb = d.get("cpu_baseline") or {}
    par = "%s reference, %s oracle" % (d.get("transcripts_identical_to_reference", "-"), d.get("transcripts_identical_to_oracle", "-"))
    print("| headline %s | %s | %s | %s (%s, %s threads) | %.2f ms | %.3f ms | %s |" % (
        d["config"]["workload"], fmt(d["value"]), fmt(d["e2e"]["value"]), fmt(cb.get("value")), cb.get("kind"), cb.get("cores"),
        d["roofline"]["kernel_ms"], d.get("roofline_prepare", {}).get("kernel_ms", float("nan")), par))
    for s in d.get("secondary", []):
        if "error" in s:
            print("| %s | error %s |" % (s["name"], s["error"]))
        elif "ms_per_call" in s:
            c = s.get("cpu_baseline") or {}
            print("| %s: %d stream(s) x %d frames per call | %.2f ms per call (median; first %.2f, last %.2f) | %s | reference %s ms per call, 1 stream | - | - | final text == reference: %s |" % (
                s["name"], s["streams"], s["chunk_frames"], s["ms_per_call"]["median"], s["ms_per_call"]["first"], s["ms_per_call"]["last"],
                fmt(s["value"]), ("%.1f" % c["ms_per_call_median"]) if c.get("ms_per_call_median") else "-", c.get("final_text_equals_b200", "-")))
        else:
            c = s.get("cpu_baseline") or {}
            print("| %s | %s | %s | %s (%s) | %.2f ms | %.3f ms | %s |" % (
                s["name"], fmt(s["value"]), fmt(s["e2e"]["value"]), fmt(c.get("value")), c.get("kind"), s["beam_kernel_ms"], s["prepare_ms"],
                s["transcripts_identical_to_oracle"]))


main()
