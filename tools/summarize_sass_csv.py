"""Summarise an `ncu --page source --csv --print-source sass` export: executed-instruction mix by opcode, warp-stall
reasons, hottest instructions.    python tools/summarize_sass_csv.py in.csv out.txt "title" """
import collections
import csv
import sys


def main():
    src, dst, title = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else sys.argv[1])
    rows = list(csv.reader(open(src)))
    kernels, cur, hdr = [], None, None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "ins": []}
            kernels.append(cur)
            hdr = None
        elif r and r[0] == "Address":
            hdr = r
            if cur is None:
                cur = {"name": "(kernel)", "ins": []}
                kernels.append(cur)
        elif hdr and cur is not None and len(r) == len(hdr):
            cur["ins"].append(dict(zip(hdr, r)))
    out = [f"# {title}", ""]
    for k in kernels:
        ins = k["ins"]
        n_exec = sum(int(i["Instructions Executed"] or 0) for i in ins)
        n_samp = sum(int(i["# Samples"] or 0) for i in ins) or 1
        out += [f"## {k['name'][:150]}", f"SASS instructions {len(ins)}, executed warp instructions {n_exec}, stall samples {n_samp}", ""]
        h, hs = collections.Counter(), collections.Counter()
        for i in ins:
            t = i["Source"].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            h[op] += int(i["Instructions Executed"] or 0)
            hs[op] += int(i["# Samples"] or 0)
        out.append("opcode        executed   share   stall-sample share")
        for op, n in h.most_common(20):
            out.append(f"{op:10s} {n:11d}  {100 * n / max(n_exec, 1):5.1f}%  {100 * hs[op] / n_samp:5.1f}%")
        st = collections.Counter()
        for i in ins:
            for key, v in i.items():
                if key.startswith("stall_") and "Not Issued" not in key and v:
                    st[key[6:]] += int(v)
        out += ["", "warp stall reasons (share of samples): " + ", ".join(f"{a}={100 * b / n_samp:.1f}%" for a, b in st.most_common(10)), ""]
        out.append("hottest instructions (stall samples | executions | SASS)")
        for i in sorted(ins, key=lambda i: -int(i["# Samples"] or 0))[:12]:
            out.append(f"{i['# Samples']:>7s} {i['Instructions Executed']:>10s}  {i['Source'][:90]}")
        out.append("")
    open(dst, "w").write("\n".join(out) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    main()
