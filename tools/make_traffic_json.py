"""profiles/traffic.json from an `ncu --page raw --csv` dump of eval_kernel launches (dram__bytes_read.sum + dram__bytes_write.sum per launch,
averaged), keyed by the hash of the kernel source the capture was made on — bench.py reports `roofline.traffic` only when the hash matches.
usage: python tools/make_traffic_json.py <eval_raw.csv> <algorithmic bytes per launch> [note]"""
import csv, hashlib, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    per = [to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]) for r in rows[2:] if "eval_kernel" in r[ik]]
    assert per, "no eval_kernel launch in the dump"
    src = os.path.join(ROOT, "featurebase_b200", "csrc", "kernels.cuh")
    out = {"eval_kernel": {"dram_bytes_per_launch": int(round(sum(per) / len(per))),
                           "kernels_cuh_sha1_16": hashlib.sha1(open(src, "rb").read()).hexdigest()[:16],
                           "source": (sys.argv[3] if len(sys.argv) > 3 else "ncu --set full, bench.py configs[1]") + ", %d launches averaged" % len(per),
                           "algorithmic_bytes_per_launch": int(sys.argv[2])}}
    json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
