"""Turns ncu CSV exports into the small summaries kept under profiles/.

  python tools/summarize_ncu.py launches <launches.csv> <out.txt> "<title>"
      <launches.csv>: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...
      -> mean device time per kernel and its share of the library's launches
  python tools/summarize_ncu.py raw <raw.csv> <out.csv> [kernel-substring]
      <raw.csv>: ncu -i capture.ncu-rep --page raw --csv
      -> one row per metric, one column per captured launch of the kernel
"""
import csv, sys, collections


def read_rows(path):
    with open(path, newline="") as f:
        rows = [r for r in csv.reader(f) if r]
    # drop ncu's "==PROF==" banner lines
    return [r for r in rows if not r[0].startswith("==")]


def launches(src, dst, title):
    rows = read_rows(src)
    head = rows[0]
    ki, mi, vi, ui = head.index("Kernel Name"), head.index("Metric Name"), head.index("Metric Value"), head.index("Metric Unit")
    acc = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
        name = r[ki].split("<")[0].replace("void ", "")
        acc.setdefault(name, []).append(us)
    ours = sum(sum(v) for k, v in acc.items() if k.startswith("pcu::"))
    with open(dst, "w") as f:
        f.write(title + "\n")
        for k, v in acc.items():
            share = 100.0 * sum(v) / ours if k.startswith("pcu::") else 0.0
            f.write("%-42s n=%3d mean %8.2f us  share %5.1f%%\n" % (k, len(v), sum(v) / len(v), share))


def raw(src, dst, needle):
    rows = read_rows(src)
    head, units, body = rows[0], rows[1], rows[2:]
    ki = head.index("Kernel Name")
    body = [r for r in body if needle in r[ki]]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(body))])
        for c, name in enumerate(head):
            if name in ("ID", "Process ID", "Process Name", "Host Name", "Context", "Stream", "Device", "CC"):
                continue
            w.writerow([name, units[c]] + [r[c] for r in body])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        raw(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
