"""Aggregate the warp-stall samples of an ncu report per CUDA source line (no GPU needed).

    python tools/ncu_lines.py <report.ncu-rep> <cubin> <kernel-name-substring> [top]

ncu's source page (csv) lists the kernel's SASS in address order; `nvdisasm -g` lists the same instructions with
`//## File ..., line N` markers (build with -lineinfo).  The two are joined by instruction index."""
import collections, csv, io, re, subprocess, sys


def sass_lines(cubin, kernel):
    txt = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
    out, cur, on = [], None, False
    for ln in txt:
        if ln.startswith(".text."):
            on = kernel in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)), "inlined" in m.group(3))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            out.append(cur)
    return out


def main():
    rep, cubin, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    import os, shlex
    # NCU_ARGS='--kernel-name regex:foo --launch-count 1' selects one launch of a multi-kernel report
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + shlex.split(os.environ.get("NCU_ARGS", "")),
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[hi + 1:]:            # the first launch only (a report may hold several)
        if r and r[0] in ("Kernel Name", "Address"):
            break
        if len(r) == len(hdr):
            data.append(r)
    lines = sass_lines(cubin, kernel)
    if len(lines) != len(data):
        print("warning: %d SASS instructions in the cubin vs %d in the report" % (len(lines), len(data)))
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot = 0
    for k, r in enumerate(data):
        key = lines[k] if k < len(lines) else None
        n = int(r[ix["# Samples"]])
        tot += n
        a = agg[key]
        a[0] += n
        a[1] += int(r[ix["Instructions Executed"]])
        for h in stalls:
            v = int(r[ix[h]])
            if v:
                a[2][h[6:]] += v
    src = {}
    print("total samples %d" % tot)
    for key, (n, ex, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if key:
            f = key[0]
            if f not in src:
                try:
                    src[f] = open("musketeer_b200/csrc/" + f).read().splitlines()
                except OSError:
                    src[f] = []
            if key[1] - 1 < len(src[f]):
                text = src[f][key[1] - 1].strip()[:80]
        print("%5.1f%% %9d  %-22s %-80s %s" % (100.0 * n / tot, ex, "%s:%d" % (key[0], key[1]) if key else "?", text,
                                               " ".join("%s:%d" % kv for kv in st.most_common(3))))


if __name__ == "__main__":
    main()
