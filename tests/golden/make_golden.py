"""Decode the reference's golden .rds vectors into a JSON fixture.

Runs only in the build container (reads /root/reference, which does not exist
on the GPU box); the JSON it writes is committed.  The .rds files are gzip'd
XDR R serialisations of a REALSXP / INTSXP with a `dim` attribute
(reference tests/testthat/test_weighted_ks.R:41-45).
"""
import gzip, json, os, struct, sys

REF = "/root/reference/tests/testthat"


def read_rds(path):
    d = gzip.open(path).read()
    assert d[:2] == b"X\n"
    off = 2 + 12  # format + 3 version ints
    flags = struct.unpack(">i", d[off:off + 4])[0]; off += 4
    typ = flags & 0xFF
    n = struct.unpack(">i", d[off:off + 4])[0]; off += 4
    if typ == 14:
        vals = list(struct.unpack(">%dd" % n, d[off:off + 8 * n])); off += 8 * n
    elif typ == 13:
        vals = list(struct.unpack(">%di" % n, d[off:off + 4 * n])); off += 4 * n
    else:
        raise ValueError(typ)
    dim = None
    if flags & 0x200:  # attribute pairlist: expect `dim`
        idx = d.find(b"dim", off)
        o2 = idx + 3
        f2 = struct.unpack(">i", d[o2:o2 + 4])[0]; o2 += 4
        assert (f2 & 0xFF) == 13
        k = struct.unpack(">i", d[o2:o2 + 4])[0]; o2 += 4
        dim = list(struct.unpack(">%di" % k, d[o2:o2 + 4 * k]))
    return {"type": "double" if typ == 14 else "integer", "dim": dim, "values": vals}


def main():
    out = {"source": "reference tests/testthat/test_weighted_ks-{es,le_prop,max_es_at}.rds",
           "recipe": "RNGversion('3.4'); set.seed(7242); scores=array(rnorm(1000*3*1), c(1000,3,1)); "
                     "gene_set=sample(seq(1000), 20)  (tests/testthat/test_weighted_ks.R:2-15)",
           "seed": 7242, "n_genes": 1000, "n_response": 3, "n_perm": 1, "set_size": 20}
    for k in ("es", "le_prop", "max_es_at"):
        out[k] = read_rds(os.path.join(REF, "test_weighted_ks-%s.rds" % k))
    # check values quoted in SURVEY.md section 4 (first scores / gene set), regenerated
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
    import flexgsea_r_b200 as fg
    rng = fg.RRNG(7242, "Rounding")
    sc = rng.rnorm(3000)
    gs = rng.sample_int(1000, 20)
    out["first_scores"] = [float(v) for v in sc[:3]]
    out["gene_set"] = [int(v) for v in gs]
    json.dump(out, open(os.path.join(os.path.dirname(__file__), "weighted_ks.json"), "w"), indent=1)
    print(json.dumps(out)[:600])


if __name__ == "__main__":
    main()
