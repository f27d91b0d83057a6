"""Freeze a handful of the reference's PUBLISHED Monte-Carlo results (data/output/*.json) as a small
fixture: tests/golden/published.json.  Run in the build container (reads /root/reference, which does not travel):

    python tests/golden/make_published.py

Each record keeps, per channel parameter, the frame count `tot`, word-error count `wec` and bit-error count `bec` the
reference's own run reached (src/main.py:37-45 stops at min_wec = 100 word errors), so that tests can put binomial
confidence intervals around the published WER / BER and hold a GPU run to them.
"""
import json
import os

REF = "/root/reference/data/output"
FILES = [
    "biawgn-1200_3_6_rand_ldpc_1-MSA-10-1.json",
    "biawgn-1200_3_6_rand_ldpc_1-SPA-10-0.json",
    "bsc-1200_3_6_rand_ldpc_1-MSA-10.json",
    "bsc-1200_3_6_rand_ldpc_1-SPA-10-0.json",
    "bec-1200_3_6_rand_ldpc_1-SPA-10-0.json",
    "bsc-1200_rho_x5_rand_ldpc_1-SPA-0-100.json",
    "biawgn-1200_rho_x5_rand_ldpc_1-SPA-0-100.json",
    "bec-1200_rho_x5_rand_ldpc_1-SPA-0-100.json",
    "biawgn-7_4_hamming-SPA-10-1.json",
    "bsc-7_4_hamming-MSA-10-1.json",
]


def main():
    out = []
    for name in FILES:
        d = json.load(open(os.path.join(REF, name)))
        n = 7 if "hamming" in d["code"] else 1200
        rec = dict(file=name, channel=d["channel"], code=d["code"], decoder=d["decoder"],
                   codeword=int(d.get("codeword", 1 if d["decoder"] == "MSA" else 0)), max_iter=int(d.get("max_iter", 10)), n=n, points=[])
        for p in d["tot"]:
            rec["points"].append(dict(param=float(p), tot=int(d["tot"][p]), wec=int(d["wec"][p]), bec=int(d["bec"][p])))
        out.append(rec)
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "published.json"), "w") as fp:
        json.dump(out, fp, indent=1)
    print("wrote", len(out), "records,", sum(len(r["points"]) for r in out), "points")


if __name__ == "__main__":
    main()
