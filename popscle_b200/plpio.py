"""dsc-pileup file formats: CEL / VAR / PLP (+ text VCF) readers and writers.

The kept input boundary of `popscle demuxlet` / `popscle freemuxlet`: the three
gzip TSV files `popscle dsc-pileup` writes (reference cmd_cram_dsc_pileup.cpp:438-523) and
`sc_dropseq_lib_t::load_from_plp` reads (sc_drop_seq.cpp:103-384), plus the genotype VCF that
demuxlet merge-joins against the VAR table (:206-330, bcf_filtered_reader.cpp:367-461).

`load_plp` reproduces the loader's semantics — droplet filters and id remap (:163-190), `--min-BQ`
/ `--cap-BQ` (:361-369), VAR/VCF merge-join with "no genotype" SNPs (:253-281), GT/GP/PL to
posterior (bcf_filtered_reader.cpp:367-461) and the genotype-error mixing (:287-315) — and returns
the flat cell-major image the C ABI takes.  Text VCF only (plain or gzip): BCF / tabix need
htslib, which this image does not have.
"""
from __future__ import annotations

import gzip
import io
from dataclasses import dataclass, field

import numpy as np

from .capi import Pileup

CEL_HEADER = "#DROPLET_ID\tBARCODE\tNUM.READ\tNUM.UMI\tNUM.UMIwSNP\tNUM.SNP"
VAR_HEADER = "#SNP_ID\tCHROM\tPOS\tREF\tALT\tAF"
PLP_HEADER = "#DROPLET_ID\tSNP_ID\tALLELES\tBASEQS"


@dataclass
class Sites:
    chrom: list          # [V] str
    pos: np.ndarray      # [V] int (1-based)
    ref: list            # [V] str (first char used)
    alt: list            # [V] str
    af: np.ndarray       # [V] float


def default_sites(V: int, af: np.ndarray, seed: int = 0) -> Sites:
    """contigs 1..22, positions strictly increasing inside a contig."""
    rng = np.random.default_rng(seed)
    per = -(-V // 22)
    chrom = [str(1 + i // per) for i in range(V)]
    pos = np.zeros(V, dtype=np.int64)
    for c in range(22):
        lo, hi = c * per, min(V, (c + 1) * per)
        if lo < hi:
            pos[lo:hi] = 10_000 + np.cumsum(rng.integers(1, 2000, hi - lo))
    bases = np.array(list("ACGT"))
    r = rng.integers(0, 4, V)
    a = (r + rng.integers(1, 4, V)) % 4
    return Sites(chrom, pos, list(bases[r]), list(bases[a]), np.asarray(af, dtype=np.float64))


def _open_w(path):
    return gzip.open(path, "wt", compresslevel=1) if str(path).endswith(".gz") else open(path, "w")


def write_bgzf(path: str, data: bytes, level: int = 1) -> None:
    """Writes `data` the way htslib's BGZF layer does (and so the way dsc-pileup writes its tables): independent gzip members of
    at most 0xff00 input bytes, each carrying its compressed size in a 'BC' extra field, closed by the 28-byte empty block."""
    import struct
    import zlib
    with open(path, "wb") as f:
        for o in range(0, len(data), 0xff00):
            chunk = data[o:o + 0xff00]
            z = zlib.compressobj(level, zlib.DEFLATED, -15)
            body = z.compress(chunk) + z.flush()
            f.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(body) + 25)
                    + body + struct.pack("<II", zlib.crc32(chunk), len(chunk)))
        f.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))


def _digits(x: np.ndarray) -> np.ndarray:
    nd = np.ones(len(x), dtype=np.int64)
    t = 10
    while len(x) and t <= int(x.max()):
        nd += x >= t
        t *= 10
    return nd


def plp_text(plp: Pileup) -> bytes:
    """The text of prefix.plp.gz for a pileup image (header + one row per pair, SNP-major, droplet id ascending inside a SNP),
    assembled with array operations: 20 M rows take seconds instead of a minute of per-row formatting."""
    C = plp.n_cells
    npair = np.diff(plp.cell_ptr)
    pair_cell = np.repeat(np.arange(C, dtype=np.int64), npair)
    order = np.lexsort((pair_cell, plp.pair_snp))
    prp = np.asarray(plp.pair_read_ptr, dtype=np.int64)
    cell, snp = pair_cell[order], np.asarray(plp.pair_snp, dtype=np.int64)[order]
    n = (prp[1:] - prp[:-1])[order]
    dc, ds = _digits(cell), _digits(snp)
    rl = dc + ds + 2 * n + 4
    off = np.concatenate(([0], np.cumsum(rl)))
    buf = np.empty(int(off[-1]), dtype=np.uint8)
    for x, nd, base in ((cell, dc, off[:-1]), (snp, ds, off[:-1] + dc + 1)):
        t = x.copy()
        for k in range(int(nd.max()) if len(nd) else 0):
            m = nd > k
            buf[(base + nd - 1 - k)[m]] = (48 + t % 10)[m].astype(np.uint8)
            t //= 10
    buf[off[:-1] + dc] = 9
    buf[off[:-1] + dc + 1 + ds] = 9
    buf[off[:-1] + dc + ds + 2 + n] = 9
    buf[off[1:] - 1] = 10
    tot = int(n.sum())
    rstart = np.concatenate(([0], np.cumsum(n)))[:-1]
    within = np.arange(tot, dtype=np.int64) - np.repeat(rstart, n)
    src = np.repeat(prp[:-1][order], n) + within
    dst = np.repeat(off[:-1] + dc + ds + 2, n) + within
    buf[dst] = np.asarray(plp.read_allele, dtype=np.uint8)[src] + 48
    buf[dst + np.repeat(n, n) + 1] = np.asarray(plp.read_qual, dtype=np.uint8)[src] + 33
    return (PLP_HEADER + "\n").encode() + buf.tobytes()


def write_plp(prefix: str, plp: Pileup, sites: Sites, barcodes=None, bgzf: bool = False):
    """Writes prefix.cel.gz / .var.gz / .plp.gz the way dsc-pileup does: PLP rows SNP-major, droplet id
    ascending inside a SNP; ALLELES '0'/'1'/'2'; BASEQS phred+33 (cmd_cram_dsc_pileup.cpp:497-518).  bgzf = as BGZF blocks
    (what htslib's writer produces) instead of one gzip stream."""
    C, V = plp.n_cells, plp.n_snps
    if barcodes is None:
        barcodes = [f"BC{c:07d}-1" for c in range(C)]
    npair = np.diff(plp.cell_ptr)
    nrd = np.diff(plp.pair_read_ptr)
    pair_cell = np.repeat(np.arange(C, dtype=np.int64), npair)
    cell_reads = np.bincount(pair_cell, weights=nrd, minlength=C).astype(np.int64)

    def emit(path, text: bytes):
        if bgzf:
            write_bgzf(path, text)
        else:
            with gzip.open(path, "wb", compresslevel=1) as f:
                f.write(text)

    emit(prefix + ".cel.gz", (CEL_HEADER + "\n" + "".join(
        f"{c}\t{barcodes[c]}\t{int(cell_reads[c])}\t{int(cell_reads[c])}\t{int(cell_reads[c])}\t{int(npair[c])}\n" for c in range(C))).encode())
    emit(prefix + ".var.gz", (VAR_HEADER + "\n" + "".join(
        f"{v}\t{sites.chrom[v]}\t{int(sites.pos[v])}\t{sites.ref[v]}\t{sites.alt[v]}\t{sites.af[v]:.5f}\n" for v in range(V))).encode())
    emit(prefix + ".plp.gz", plp_text(plp))
    return barcodes


def write_vcf(path: str, sites: Sites, samples, geno=None, gp=None, pl=None, keep=None, r2=None):
    """Text VCF with GT (geno int [nv][V], -1 = missing), GP (float [V][nv][3]) and/or PL
    (int [V][nv][3]) FORMAT fields.  keep = mask of SNPs to emit (the rest are absent from the VCF)."""
    V = len(sites.chrom)
    contigs = []
    for c in sites.chrom:
        if not contigs or contigs[-1] != c:
            if c not in contigs:
                contigs.append(c)
    fmt = [k for k, x in (("GT", geno), ("GP", gp), ("PL", pl)) if x is not None]
    with _open_w(path) as f:
        f.write("##fileformat=VCFv4.2\n")
        for c in contigs:
            f.write(f"##contig=<ID={c}>\n")
        f.write('##INFO=<ID=AF,Number=A,Type=Float,Description="Allele Frequency">\n')
        f.write('##INFO=<ID=R2,Number=1,Type=Float,Description="Imputation quality">\n')
        f.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
        f.write('##FORMAT=<ID=GP,Number=G,Type=Float,Description="Genotype posterior">\n')
        f.write('##FORMAT=<ID=PL,Number=G,Type=Integer,Description="Phred-scaled genotype likelihood">\n')
        f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(samples) + "\n")
        gts = {-1: "./.", 0: "0/0", 1: "0/1", 2: "1/1"}
        for v in range(V):
            if keep is not None and not keep[v]:
                continue
            cols = []
            for j in range(len(samples)):
                parts = []
                if geno is not None:
                    parts.append(gts[int(geno[j, v])])
                if gp is not None:
                    parts.append(",".join(f"{x:.4g}" for x in gp[v, j]))
                if pl is not None:
                    parts.append(",".join(str(int(x)) for x in pl[v, j]))
                cols.append(":".join(parts))
            f.write(f"{sites.chrom[v]}\t{int(sites.pos[v])}\t.\t{sites.ref[v]}\t{sites.alt[v]}\t.\tPASS\tAF={sites.af[v]:.5f};R2={0.9 if r2 is None else float(r2[v]):.4g}\t"
                    + ":".join(fmt) + "\t" + "\t".join(cols) + "\n")


# ---------------------------------------------------------------------------------------------------
# readers
# ---------------------------------------------------------------------------------------------------
def _rows(path):
    """whitespace-split rows of a plain / gzip text file; stops at the first empty line like
    tsv_reader::read_line (tsv_reader.cpp:37-41 returns 0 for lstr <= 0)"""
    with open(path, "rb") as fh:
        gz = fh.read(2) == b"\x1f\x8b"
    f = gzip.open(path, "rt") if gz else open(path, "r")
    with f:
        for line in f:
            line = line.rstrip("\n")
            if not line:
                return
            yield line.split()


@dataclass
class Genotypes:
    samples: list
    gp: np.ndarray       # float64 [V][nv][3] after the error mixing (what add_snp stores)
    has_gp: np.ndarray   # uint8 [V]
    # raw forms of ABI 4 (what the hosts hand to the library, which mixes on the device)
    gp_f32: np.ndarray | None = None   # float32 [V][nv][3] as the reader returned them
    err_snp: np.ndarray | None = None  # float64 [V] error rate before clamping
    gt8: np.ndarray | None = None      # uint8 [V][nv] hard calls when every row with GP is one-hot, else None

    def raw(self):
        """RawGeno for Context.demux_run: hard calls when possible, float posteriors otherwise."""
        from .capi import RawGeno
        if self.gt8 is not None:
            return RawGeno(gt8=self.gt8, err_snp=self.err_snp, err=0.0)
        return RawGeno(gp_f32=self.gp_f32, err_snp=self.err_snp, err=0.0)


@dataclass
class LoadedPileup:
    plp: Pileup
    barcodes: list            # compact cell id -> barcode
    sites: Sites
    cell_uniq_reads: np.ndarray  # NUM.READS column of .best (cell_uniq_reads, cmd_cram_demuxlet.cpp:997)
    cell_totl_reads: np.ndarray  # what --min-total is compared with (cmd_cram_demuxlet.cpp:641)
    geno: Genotypes | None = None
    rid2chr: list = field(default_factory=list)


def _phred2prob(q):
    return 10.0 ** (-0.1 * np.asarray(q, dtype=np.float64))  # phredConv.toProb


def _parse_vcf(path, field, sm_list=None, min_mac=1, min_callrate=0.5, max_alleles=2):
    """Records of a text VCF that pass the demuxlet site filters (cmd_cram_demuxlet.cpp:27-29,
    bcf_filtered_reader.cpp:505-581), with per-sample float32 posteriors
    (bcf_filtered_reader.cpp:367-461, gt_error = 0 as load_from_plp passes).  Yields
    (contig rid, pos, ref, alt, gp[nv*3] float32, INFO column)."""
    contigs, samples, cols = {}, None, None
    for line in _lines(path):
        if line.startswith("##"):
            if line.startswith("##contig=<ID="):
                cid = line[len("##contig=<ID="):].split(",")[0].rstrip(">")
                contigs.setdefault(cid, len(contigs))
            continue
        if line.startswith("#CHROM"):
            samples = line.split("\t")[9:]
            if sm_list:  # std::set order = sorted (bcf_filtered_reader.cpp:107-125)
                names = sorted(set(sm_list))
                cols = [samples.index(s) for s in names]
                use = names
            else:
                cols = list(range(len(samples)))
                use = samples
            yield ("header", use, contigs)
            continue
        c = line.split("\t")
        rid = contigs.setdefault(c[0], len(contigs))
        pos = int(c[1])
        alleles = [c[3]] + ([] if c[4] == "." else c[4].split(","))
        nal = len(alleles)
        fmt = c[8].split(":")
        smp = [x.split(":") for x in c[9:]]
        if nal > max_alleles:
            continue
        # GT is parsed (and required) only when something reads it: the site filters minMAC / minCallRate
        # (require_GT, bcf_filter_arg.h:110-113) or --field GT itself
        need_gt = min_mac > 0 or min_callrate > 0 or field == "GT"
        if "GT" not in fmt and need_gt:
            raise ValueError(f"Cannot find the field GT from the VCF file at position {c[0]}:{pos}")
        gts = [(-1, -1)] * len(cols)
        if "GT" in fmt:
            gi = fmt.index("GT")
            gts = []
            for j in cols:
                a = smp[j][gi].replace("|", "/").split("/")
                a = [(-1 if x in (".", "") else int(x)) for x in a] + [-1]
                if a[0] >= nal or a[1] >= nal:
                    raise ValueError(f"GT {smp[j][gi]} names an allele beyond the ALT list at position {c[0]}:{pos}")
                gts.append((a[0], a[1]))
        an = sum((x >= 0) + (y >= 0) for x, y in gts)
        acs = [0] * nal
        for x, y in gts:
            if x >= 0: acs[x] += 1
            if y >= 0: acs[y] += 1
        nv = len(cols)
        if need_gt:
            if nv and min_callrate > an / (2.0 * nv):
                continue
            ac = an - acs[0]
            if ac < min_mac or an - ac < min_mac:
                continue
        ngen = nal * (nal + 1) // 2
        gp = np.zeros(nv * ngen, dtype=np.float32)
        if field == "GT":
            for i, (x, y) in enumerate(gts):
                if x < 0 or y < 0:  # missing: HWE from smoothed counts (:385-397)
                    l = 0
                    for j in range(nal):
                        for k in range(j + 1):
                            gp[i * ngen + l] = np.float32((1.0 if j == k else 2.0) * (acs[j] + 1.0 / nal) / (an + 1.0) * (acs[k] + 1.0 / nal) / (an + 1.0))
                            l += 1
                else:
                    a, b = max(x, y), min(x, y)
                    g = a * (a + 1) // 2 + b
                    gp[i * ngen: (i + 1) * ngen] = 0.0  # gt_error = 0: (g == j) ? 1 : 0/(ngenos-1)
                    gp[i * ngen + g] = 1.0
        elif field == "PL":
            k = fmt.index("PL")
            pls = np.array([[int(x) for x in smp[j][k].split(",")] for j in cols], dtype=np.float64)
            af = np.full(nal, 1.0 / nal)
            post = None
            for _ in range(10):  # :267-318 (ploidy 2 only)
                new = np.zeros(nal)
                post = np.zeros((nv, ngen))
                for i in range(nv):
                    l = 0
                    for j in range(nal):
                        for kk in range(j + 1):
                            post[i, l] = (1 if j == kk else 2) * af[j] * af[kk] * _phred2prob(pls[i, l])
                            l += 1
                    post[i] /= post[i].sum()
                    l = 0
                    for j in range(nal):
                        for kk in range(j + 1):
                            new[j] += post[i, l]; new[kk] += post[i, l]
                            l += 1
                af = new / (2 * nv)
            gp = post.astype(np.float32).ravel()
        else:  # GP-like float field (:434-458), gt_error = 0
            k = fmt.index(field)
            raw = np.array([[float(x) for x in smp[j][k].split(",")] for j in cols], dtype=np.float32)
            s = np.zeros(nv, dtype=np.float32)
            for g in range(ngen):
                s = (s + raw[:, g]).astype(np.float32)
            gp = (raw / s[:, None]).astype(np.float32).ravel()
        yield (rid, pos, alleles[0], alleles[1] if nal > 1 else ".", gp, c[7])


def _lines(path):
    with open(path, "rb") as fh:
        gz = fh.read(2) == b"\x1f\x8b"
    f = gzip.open(path, "rt") if gz else open(path, "r")
    with f:
        for line in f:
            line = line.rstrip("\r\n")
            if line:
                yield line


def _info_float(info: str, key: str):
    """One float32 value of an INFO key (bcf_get_info_float; the reference insists on exactly one value)."""
    for kv in info.split(";"):
        if kv.startswith(key + "="):
            vals = kv[len(key) + 1:].split(",")
            if len(vals) != 1 or vals[0] in (".", ""):
                return None
            return float(np.float32(float(vals[0])))
    return None


def load_plp(prefix: str, vcf: str | None = None, field: str = "GP", geno_error_offset: float = 0.1,
             geno_error_coeff: float = 0.0, r2_info: str = "R2", sm_list=None, min_bq: int = 1, cap_bq: int = 60, min_read: int = 0, min_umi: int = 0, min_snp: int = 0,
             group_list=None, min_mac: int = 1, min_callrate: float = 0.5) -> LoadedPileup:
    """sc_dropseq_lib_t::load_from_plp (sc_drop_seq.cpp:103-384).  Library defaults min_bq=1 /
    cap_bq=60 (sc_drop_seq.h:181); the demuxlet and freemuxlet commands pass 13 / 20."""
    # ---- CEL ----
    rows = _rows(prefix + ".cel.gz")
    hdr = next(rows, None)
    if hdr is None or hdr != CEL_HEADER.split("\t"):
        raise ValueError(f"The header line of {prefix}.cel.gz is malformed or outdated. Expecting " + CEL_HEADER.replace("\t", " "))
    valid = set(group_list) if group_list else None
    index_bcs, barcodes, tmp_totl, tmp_uniq, tmp_nsnp = [], [], [], [], []
    nskip = 0
    for r in rows:
        if valid is not None and r[1] not in valid:
            nskip += 1; index_bcs.append(-1); continue
        n_reads, n_umis, n_uws, n_snps = (int(_atoi(x)) for x in r[2:6])
        if n_reads < min_read or n_umis < min_umi or n_snps < min_snp:
            index_bcs.append(-1); nskip += 1; continue
        new_id = len(barcodes)
        if new_id + nskip != int(r[0]):
            raise ValueError(f"Observed DROPLET_ID {r[0]} is different from expected DROPLET_ID. Did you modify the digital pileup files by yourself?")
        barcodes.append(r[1]); index_bcs.append(new_id)
        tmp_totl.append(n_reads); tmp_uniq.append(n_uws); tmp_nsnp.append(n_snps)
    index_bcs = np.asarray(index_bcs, dtype=np.int64)
    C = len(barcodes)
    # ---- VAR (+ VCF merge-join) ----
    rows = _rows(prefix + ".var.gz")
    hdr = next(rows, None)
    if hdr is None or hdr != VAR_HEADER.split("\t"):
        raise ValueError(f"The header line of {prefix}.var.gz is malformed or outdated. Expecting " + VAR_HEADER.replace("\t", " "))
    chr2rid, rid2chr = {}, []
    chrom, pos, ref, alt, af = [], [], [], [], []
    vit = samples = cur = None
    gps, has, raws, errs = [], [], [], []
    if vcf is not None:
        vit = _parse_vcf(vcf, field, sm_list, min_mac, min_callrate)
        _, samples, _ = next(vit)
        cur = next(vit, None)
        if cur is None:
            raise ValueError(f"Cannot read any single variant from {vcf}")
    nv = len(samples) if samples is not None else 0
    for n, r in enumerate(rows):
        if r[1] not in chr2rid:
            chr2rid[r[1]] = len(chr2rid); rid2chr.append(r[1])
        rid, p = chr2rid[r[1]], int(_atoi(r[2]))
        if int(_atoi(r[0])) != n:
            pass  # the reference checks the running count, not the id column (:240)
        chrom.append(r[1]); pos.append(p); ref.append(r[3][0]); alt.append(r[4][0]); af.append(float(r[5]))
        if vit is None:
            continue
        found = False
        while True:  # :253-281 — VCF rids are compared NUMERICALLY with the VAR first-appearance rids
            if cur is None or cur[0] > rid:
                break
            if cur[0] == rid:
                if cur[1] > p:
                    break
                if cur[1] == p:
                    found = (cur[2][0] == r[3][0]) and (cur[3][0] == r[4][0])
                    break
            cur = next(vit, None)
        if not found:
            has.append(0); gps.append(np.zeros(nv * 3)); raws.append(np.zeros(nv * 3, dtype=np.float32)); errs.append(0.0); continue
        raws.append(np.asarray(cur[4], dtype=np.float32)[:nv * 3])
        g = cur[4].astype(np.float64)  # get_posterior_at widens the float (bcf_filtered_reader.h:166)
        avg = np.full(3, 1e-10)
        for i in range(nv * 3):  # :289-291, sample-major accumulation order
            avg[i % 3] += g[i]
        avg = avg / (avg[0] + avg[1] + avg[2])
        err = geno_error_offset
        if geno_error_coeff > 0:  # [error] = [offset] + [1-offset]*[1-R2]*[coeff] (sc_drop_seq.cpp:299-306)
            r2 = _info_float(cur[5], r2_info)
            if r2 is None:
                raise ValueError(f"Cannot extract {r2_info} (1 float value) from INFO field at {r[1]}:{p}. Cannot use --geno-error-coeff")
            # `1 - r2flts[0]` is evaluated in float32 in the reference (int - float), then widened
            err += (1 - geno_error_offset) * float(np.float32(1.0) - np.float32(r2)) * geno_error_coeff
        errs.append(err)
        err = min(max(err, 0.0), 0.999)
        if err > 0:
            g = (1 - err) * g + err * np.tile(avg, nv)
        has.append(1); gps.append(g)
    V = len(chrom)
    sites = Sites(chrom, np.asarray(pos, dtype=np.int64), ref, alt, np.asarray(af, dtype=np.float64))
    # ---- PLP ----
    rows = _rows(prefix + ".plp.gz")
    hdr = next(rows, None)
    if hdr is None or hdr != PLP_HEADER.split("\t"):
        raise ValueError(f"The header line of {prefix}.plp.gz is malformed or outdated. Expecting " + PLP_HEADER.replace("\t", " "))
    rc, rs, ra, rq, rn = [], [], [], [], []
    for r in rows:
        ibc = index_bcs[int(r[0])]
        if ibc < 0:
            continue
        q = np.frombuffer(r[3].encode(), dtype=np.uint8).astype(np.int16) - 33
        a = np.frombuffer(r[2].encode(), dtype=np.uint8).astype(np.int16)[:len(q)] - ord("0")
        keep = q >= min_bq
        if not keep.any():
            continue
        rc.append(ibc); rs.append(int(r[1])); rn.append(int(keep.sum()))
        ra.append(a[keep].astype(np.uint8)); rq.append(np.minimum(q[keep], cap_bq).astype(np.uint8))
    rc = np.asarray(rc, dtype=np.int64); rs = np.asarray(rs, dtype=np.int64); rn = np.asarray(rn, dtype=np.int64)
    # cell-major, SNP ascending; a (cell,SNP) listed on several rows keeps its rows in file order
    order = np.lexsort((np.arange(len(rc)), rs, rc))
    key = rc[order] * max(V, 1) + rs[order]
    first = np.concatenate([[True], key[1:] != key[:-1]]) if len(key) else np.zeros(0, bool)
    pair_of_row = np.cumsum(first) - 1
    P = int(first.sum())
    pair_cell = rc[order][first]
    pair_snp = rs[order][first].astype(np.int32)
    nrd = np.bincount(pair_of_row, weights=rn[order], minlength=P).astype(np.int64) if P else np.zeros(0, np.int64)
    prp = np.concatenate([[0], np.cumsum(nrd)]).astype(np.int64)
    al = np.concatenate([ra[i] for i in order]) if len(order) else np.zeros(0, np.uint8)
    bq = np.concatenate([rq[i] for i in order]) if len(order) else np.zeros(0, np.uint8)
    cell_ptr = np.concatenate([[0], np.cumsum(np.bincount(pair_cell, minlength=C))]).astype(np.int64)
    plp = Pileup(C, V, cell_ptr, pair_snp, prp, al, bq, sites.af)
    uniq = np.bincount(pair_cell, weights=nrd, minlength=C).astype(np.int64) if P else np.zeros(C, np.int64)
    # sanity check on the observed counts (:375-381): NUM.READ of the CEL file replaces the pass count
    # only where NUM.UMIwSNP and NUM.SNP agree with what was loaded
    npair_cell = np.diff(cell_ptr)
    totl = uniq.copy()
    agree = (uniq == np.asarray(tmp_uniq, dtype=np.int64)) & (npair_cell == np.asarray(tmp_nsnp, dtype=np.int64)) if C else np.zeros(0, bool)
    totl[agree] = np.asarray(tmp_totl, dtype=np.int64)[agree]
    geno = None
    if vit is not None:
        has8 = np.asarray(has, dtype=np.uint8)
        f32 = np.asarray(raws, dtype=np.float32).reshape(V, nv, 3)
        onehot = ((f32 == 1.0).sum(axis=2) == 1) & ((f32 == 0.0).sum(axis=2) == 2)
        gt8 = None
        if bool(onehot[has8 != 0].all()):  # --field GT without missing calls: one byte per call crosses PCIe
            gt8 = np.where(has8[:, None] != 0, f32.argmax(axis=2), 0).astype(np.uint8)
        geno = Genotypes(list(samples), np.asarray(gps, dtype=np.float64).reshape(V, nv, 3), has8,
                         np.ascontiguousarray(f32), np.asarray(errs, dtype=np.float64), gt8)
    return LoadedPileup(plp, barcodes, sites, uniq, totl, geno, rid2chr)


def _atoi(s: str) -> int:
    """C atoi: leading integer or 0 ('.' in NUM.UMI with --skip-umi, cmd_cram_dsc_pileup.cpp:459)"""
    n, i = 0, 0
    neg = s.startswith("-")
    if neg or s.startswith("+"):
        i = 1
    while i < len(s) and s[i].isdigit():
        n = n * 10 + ord(s[i]) - 48
        i += 1
    return -n if neg else n
