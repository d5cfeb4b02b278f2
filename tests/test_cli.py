"""The `fulgor_b200_pseudoalign` CLI (fulgor_b200/csrc/pseudoalign_cli.cpp) against the oracle and, where its binary was
built (oracle/_ref/fulgor_ref, compiled from the unmodified reference sources), against the reference's own
`fulgor pseudoalign` run on the same files: same records, compared after sorting by read id (the reference's record
order depends on thread scheduling, reference README.md:220)."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import _checkers as ck

GPU_CLI = os.path.join(ck.ROOT, "fulgor_b200", "fulgor_b200_pseudoalign")
HOST_CLI = os.path.join(ck.ROOT, "build", "fulgor_cli_hosttest")


def _build_host_cli():
    """build/fulgor_cli_hosttest = the tool's host code (fulgor_b200/csrc/pseudoalign_cli.cpp) linked against
    tests/fake_gpu_lib.cpp, a test double of the C ABI answered by the oracle"""
    ck.build_checkers()
    srcs = [os.path.join(ck.ROOT, "fulgor_b200", "csrc", "pseudoalign_cli.cpp"), os.path.join(ck.ROOT, "tests", "fake_gpu_lib.cpp")]
    deps = srcs + [os.path.join(ck.ROOT, "fulgor_b200", "csrc", "fastx_io.h"), os.path.join(ck.ROOT, "include", "fulgor_gpu.h"), ck.ORACLE_SO]
    if not os.path.exists(HOST_CLI) or any(os.path.getmtime(d) > os.path.getmtime(HOST_CLI) for d in deps):
        os.makedirs(os.path.dirname(HOST_CLI), exist_ok=True)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-w", "-o", HOST_CLI] + srcs +
                              ["-L" + os.path.dirname(ck.ORACLE_SO), "-lfulgor_oracle", "-Wl,-rpath," + os.path.dirname(ck.ORACLE_SO), "-lz"])
    return HOST_CLI


@pytest.fixture(scope="module", params=[pytest.param("gpu", marks=pytest.mark.gpu), "host"])
def cli(request, built_lib):
    """the tool under test: "gpu" = the shipped binary on a B200 (GPU tier); "host" = the SAME host code linked against the
    oracle-backed test double of the C ABI, so that the CPU tier covers the tool's host logic: arguments, feeder, batch
    pipeline, multi-device split and splice, de-duplication fan-out, output formats, the per-k-mer tools' lines.
    Returns (path, scale of the read counts: the oracle behind the double is single-threaded)."""
    if request.param == "gpu":
        return GPU_CLI, 1.0
    return _build_host_cli(), 0.3


def write_fastq(path, reads, gz=False, fasta=False):
    bases, off = reads
    op = gzip.open if gz else open
    with op(path, "wb") as f:
        for i in range(len(off) - 1):
            s = bases[int(off[i]):int(off[i + 1])].tobytes()
            if fasta:
                f.write(b">r%d\n%s\n" % (i, s))
            else:
                f.write(b"@r%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)))


def ascii_records(path):
    recs = {}
    for line in open(path, "rb").read().split(b"\n"):
        if not line:
            continue
        f = line.split(b"\t")
        rid, n = int(f[0]), int(f[1])
        assert len(f) == 2 + n
        recs[rid] = [int(x) for x in f[2:]]
    return recs


def binary_records(path):
    a = np.fromfile(path, dtype=np.uint32)
    recs, p = {}, 0
    while p < a.size:
        rid, n = int(a[p]), int(a[p + 1])
        recs[rid] = a[p + 2: p + 2 + n].tolist()
        p += 2 + n
    return recs


class BitReader:
    def __init__(self, words, nbits):
        self.w, self.n, self.p = words, nbits, 0

    def take(self, l):
        v = 0
        for i in range(l):
            v |= ((int(self.w[(self.p + i) >> 6]) >> ((self.p + i) & 63)) & 1) << i
        self.p += l
        return v

    def unary(self):
        z = 0
        while self.take(1) == 0:
            z += 1
        return z

    def gamma(self):
        b = self.unary()
        return (self.take(b) | (1 << b)) - 1

    def delta(self):
        b = self.gamma()
        return (self.take(b) | (1 << b)) - 1


def compressed_records(path):
    """decoder for psa_compressed_formatter's output (reference src/ps_utils.cpp:149-243)"""
    raw = open(path, "rb").read()
    C = int.from_bytes(raw[:8], "little")
    sparse, dense = int(0.25 * C), int(0.75 * C)
    recs, p = {}, 8
    while p < len(raw):
        nbits = int.from_bytes(raw[p:p + 8], "little")
        nwords = (nbits + 63) // 64
        words = np.frombuffer(raw, dtype="<u8", count=nwords, offset=p + 8)
        p += 8 + 8 * nwords
        r = BitReader(words, nbits)
        while r.p < nbits:
            rid, size = r.delta(), r.delta()
            if size == 0:
                vals = []
            elif size < sparse:
                vals = [r.delta()]
                for _ in range(size - 1):
                    vals.append(vals[-1] + r.delta() + 1)
            elif size < dense:
                bits = r.take(C)
                vals = [c for c in range(C) if (bits >> c) & 1]
            else:
                missing = []
                for i in range(C - size):
                    missing.append(r.delta() if i == 0 else missing[-1] + r.delta() + 1)
                ms = set(missing)
                vals = [c for c in range(C) if c not in ms]
            recs[rid] = vals
    return C, recs


def csr_records(csr):
    off, vals = csr
    return {i: vals[int(off[i]):int(off[i + 1])].tolist() for i in range(len(off) - 1)}


def test_cli_writes_a_pipe_like_a_file(cli, tmp_path):
    """the formatting threads place their pieces with positional writes; an output that cannot seek (a pipe: `-o /dev/stdout |`)
    takes the pieces in order instead -- same bytes, with several threads and several batches, in every format"""
    CLI, scale = cli
    reads = ck.gen_reads(int(6000 * scale), 75, 300, seed=41, genomes="salmonella_10")
    fq = str(tmp_path / "reads.fq")
    write_fastq(fq, reads)
    path = ck.index_path("salmonella_10.fur")
    for fmt in ("ascii", "binary", "compressed"):
        out = str(tmp_path / f"out.{fmt}")
        args = [CLI, "-i", path, "-q", fq, "-t", "6", "--format", fmt, "--batch-reads", str(int(2500 * scale))]
        subprocess.check_call(args + ["-o", out])
        piped = subprocess.run(args + ["-o", "/dev/stdout"], stdout=subprocess.PIPE, check=True).stdout
        assert piped == open(out, "rb").read() and len(piped) > 0


@pytest.mark.parametrize("index,algo_args", [("salmonella_10.fur", []), ("salmonella_10.fur", ["-r", "0.8"]), ("salmonella_10.mfur", []),
                                             ("synth_200.fur", []), ("synth_200.mfur", ["-r", "0.6"]),
                                             ("salmonella_10.dfur", ["-r", "0.7"]), ("salmonella_10.mdfur", []),
                                             ("synth_200.dfur", []), ("synth_200.mdfur", ["-r", "0.5"])])
def test_cli_matches_oracle_and_reference(index, algo_args, cli, tmp_path):
    CLI, scale = cli
    genomes = index.split(".")[0]
    reads = ck.gen_reads(int(3000 * scale), 75, 300, seed=31, genomes=genomes)
    fq = str(tmp_path / "reads.fq")
    write_fastq(fq, reads)
    path = ck.index_path(index)
    o = ck.Oracle(path)
    thr = float(algo_args[1]) if algo_args else 1.0
    exp = csr_records(o.pseudoalign(reads, 1 if algo_args else 0, thr))
    outs = {}
    for fmt in ("ascii", "binary", "compressed"):
        out = str(tmp_path / f"out.{fmt}")
        subprocess.check_call([CLI, "-i", path, "-q", fq, "-o", out, "--format", fmt, "--batch-reads", str(int(1000 * scale))] + algo_args)
        outs[fmt] = out
    assert ascii_records(outs["ascii"]) == exp
    assert binary_records(outs["binary"]) == exp
    C, recs = compressed_records(outs["compressed"])
    assert C == o.num_colors and recs == exp
    # exact text of the ascii format: "id \t n [\t c]* \n" in read order
    want = b"".join(b"\t".join([b"%d" % i, b"%d" % len(exp[i])] + [b"%d" % c for c in exp[i]]) + b"\n" for i in range(len(exp)))
    assert open(outs["ascii"], "rb").read() == want
    if os.path.exists(ck.REF_CLI):
        for fmt, parse in (("ascii", ascii_records), ("binary", binary_records)):
            ref_out = str(tmp_path / f"ref.{fmt}")
            subprocess.check_call([ck.REF_CLI, "pseudoalign", "-i", path, "-q", fq, "-o", ref_out, "-t", "4", "--format", fmt] + algo_args,
                                  stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            assert parse(ref_out) == parse(outs[fmt])
        ref_out = str(tmp_path / "ref.compressed")
        subprocess.check_call([ck.REF_CLI, "pseudoalign", "-i", path, "-q", fq, "-o", ref_out, "-t", "2", "--format", "compressed"] + algo_args,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        assert compressed_records(ref_out) == (C, recs)


@pytest.mark.parametrize("index", ["salmonella_10.fur", "synth_200.mfur", "synth_200.dfur"])
def test_cli_deduplicate(index, cli, tmp_path):
    """--deduplicate (tools/pseudoalign.cpp:92-226): same records as without it, and as the reference's own --deduplicate run"""
    CLI, scale = cli
    genomes = index.split(".")[0]
    base = ck.gen_reads(500, 75, 300, seed=41, genomes=genomes)
    seqs = [base[0][int(base[1][i]):int(base[1][i + 1])].tobytes() for i in range(500)]
    rng = np.random.default_rng(9)
    reads = ck.reads_from_list([seqs[j] for j in rng.integers(0, 500, int(4000 * scale))] + [b"ACGT", b"N" * 100])
    fq = str(tmp_path / "reads.fq")
    write_fastq(fq, reads)
    path = ck.index_path(index)
    exp = csr_records(ck.Oracle(path).pseudoalign(reads, 0))
    for fmt, parse in (("ascii", ascii_records), ("binary", binary_records)):
        out = str(tmp_path / f"out.{fmt}")
        subprocess.check_call([CLI, "-i", path, "-q", fq, "-o", out, "--format", fmt, "--batch-reads", str(int(1500 * scale)), "--deduplicate"])
        assert parse(out) == exp
    if os.path.exists(ck.REF_CLI):
        ref_out = str(tmp_path / "ref.ascii")
        subprocess.check_call([ck.REF_CLI, "pseudoalign", "-i", path, "-q", fq, "-o", ref_out, "-t", "4", "--deduplicate"], cwd=str(tmp_path),
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        got = ascii_records(ref_out)
        # The reference's duplicate-marking loop starts with `next = curr++` (tools/pseudoalign.cpp:196-197), so when the two
        # lexicographically smallest lists are equal BOTH lose their list and are written with an empty result (which two
        # reads of that group depends on its thread scheduling). Everything else must match; those two must be empty.
        bad = [i for i in exp if got.get(i) != exp[i]]
        assert len(got) == len(exp) and len(bad) <= 2 and all(got[i] == [] for i in bad)
        if bad:
            cid_off, cids = ck.Oracle(path).fetch_color_set_ids(reads)
            lists = [tuple(cids[int(cid_off[i]):int(cid_off[i + 1])].tolist()) for i in range(len(cid_off) - 1)]
            smallest = min(l for l in lists if l)
            assert len(bad) == 2 and all(lists[i] == smallest for i in bad)
    r = subprocess.run([CLI, "-i", path, "-q", fq, "-o", str(tmp_path / "x"), "--deduplicate", "-r", "0.5"], capture_output=True, text=True)
    assert r.returncode == 1 and "Deduplication not available" in r.stderr


@pytest.mark.parametrize("index", ["salmonella_10.fur", "synth_200.mdfur"])
def test_cli_kmer_tools(index, cli, tmp_path):
    """`kmer-conservation` / `kmer-matches` front-ends: the reference's output lines (tools/kmer_conservation.cpp:27-37,
    tools/kmer_matches.cpp:28-34), checked against the oracle and, where it was built, against the reference's own tools
    (their line order depends on thread scheduling: compared as sorted lines)"""
    CLI, scale = cli
    genomes = index.split(".")[0]
    reads = ck.gen_reads(600, 75, 300, seed=51, genomes=genomes)
    fq = str(tmp_path / "reads.fq")
    write_fastq(fq, reads)
    path = ck.index_path(index)
    o = ck.Oracle(path)
    n = len(reads[1]) - 1
    out_c, out_m = str(tmp_path / "cons.txt"), str(tmp_path / "match.txt")
    subprocess.check_call([CLI, "kmer-conservation", "-i", path, "-q", fq, "-o", out_c, "--batch-reads", "250"])
    subprocess.check_call([CLI, "kmer-matches", "-i", path, "-q", fq, "-o", out_m, "--batch-reads", "250"])
    if scale < 1:  # host tier: also the multi-threaded formatting of one large batch
        for tool, first in (("kmer-conservation", out_c), ("kmer-matches", out_m)):
            again = str(tmp_path / (tool + ".t6"))
            subprocess.check_call([CLI, tool, "-i", path, "-q", fq, "-o", again, "-t", "6"])
            assert open(again, "rb").read() == open(first, "rb").read()
    toff, tr = o.kmer_conservation(reads)
    want = []
    for i in range(n):
        t = tr[int(toff[i]):int(toff[i + 1])]
        want.append("\t".join(["r%d" % i, str(len(t))] + ["(%d %d %d)" % tuple(x) for x in t]))
    assert open(out_c).read().split("\n")[:-1] == want
    koff, pos, counts = o.kmer_matches(reads)
    want = []
    for i in range(n):
        p = pos[int(koff[i]):int(koff[i + 1])]
        want.append("\t".join(["r%d" % i, str(len(p))] + [str(int(b)) for b in p] + [str(int(c)) for c in counts[i]]))
    assert open(out_m).read().split("\n")[:-1] == ["num_colors=%d" % o.num_colors] + want
    if os.path.exists(ck.REF_CLI):
        for tool, mine in (("kmer-conservation", out_c), ("kmer-matches", out_m)):
            ref_out = str(tmp_path / ("ref_" + tool))
            r = subprocess.run([ck.REF_CLI, tool, "-i", path, "-q", fq, "-o", ref_out, "-t", "3"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            if r.returncode != 0:
                pytest.skip("this build of oracle/_ref/fulgor_ref has no " + tool)
            ref_lines, my_lines = sorted(open(ref_out).read().split("\n")), sorted(open(mine).read().split("\n"))
            if tool == "kmer-conservation":
                assert ref_lines == my_lines
                continue
            # kmer-matches: each worker of the reference tool reuses one bit_vector::builder across its reads and builder::resize
            # keeps the old words (bits/include/bit_vector.hpp:27-36), so its positive-k-mer column also shows stale ones from
            # the worker's earlier reads (which depends on thread scheduling). Names, k-mer counts and per-color counts must be
            # identical; the reference's bits must be a superset of ours (ours == index::kmer_matches on a fresh builder, checked
            # against the oracle above).
            assert len(ref_lines) == len(my_lines)
            for a, b in zip(ref_lines, my_lines):
                fa, fb = a.split("\t"), b.split("\t")
                if len(fa) < 2:
                    assert a == b
                    continue
                nk = int(fa[1])
                assert fa[:2] == fb[:2] and fa[2 + nk:] == fb[2 + nk:]
                assert all(x == y or (x == "1" and y == "0") for x, y in zip(fa[2:2 + nk], fb[2:2 + nk]))


def test_cli_gz_fasta_and_verbose_summary(cli, tmp_path):
    CLI, scale = cli
    reads = ck.gen_reads(2000, seed=5)
    fa = str(tmp_path / "reads.fa.gz")
    write_fastq(fa, reads, gz=True, fasta=True)
    path = ck.index_path("salmonella_10.fur")
    out = str(tmp_path / "out.txt")
    p = subprocess.run([CLI, "-i", path, "-q", fa, "-o", out, "--verbose"], capture_output=True, text=True, check=True)
    exp = csr_records(ck.Oracle(path).pseudoalign(reads, 0))
    assert ascii_records(out) == exp
    mapped = sum(1 for v in exp.values() if v)
    assert "processed 2000 reads" in p.stdout and "musec/read" in p.stdout
    assert f"num_mapped_reads {mapped}/2000" in p.stdout


def test_cli_multi_device_split_and_splice(tmp_path):
    """host tier only (the double offers two "devices"): --gpus 2 cuts every batch by cumulative k-mer count, runs the halves on
    two handles from two threads and splices the CSR results back in read order; with --deduplicate the representatives'
    indexes are rebased onto the batch"""
    _build_host_cli()
    reads = ck.gen_reads(900, 75, 300, seed=61, genomes="synth_200")
    base = [reads[0][int(reads[1][i]):int(reads[1][i + 1])].tobytes() for i in range(900)]
    reads = ck.reads_from_list(base + base[:300] + [b"", b"ACGT"])
    fq = str(tmp_path / "reads.fq")
    write_fastq(fq, reads)
    path = ck.index_path("synth_200.mfur")
    o = ck.Oracle(path)
    for args, algo, thr in ((["--gpus", "2"], 0, 1.0), (["--gpus", "2", "-r", "0.6"], 1, 0.6), (["--gpus", "2", "--deduplicate", "-t", "6"], 0, 1.0)):
        out = str(tmp_path / "out.txt")
        subprocess.check_call([HOST_CLI, "-i", path, "-q", fq, "-o", out, "--batch-reads", "500"] + args)
        assert ascii_records(out) == csr_records(o.pseudoalign(reads, algo, thr))
    r = subprocess.run([HOST_CLI, "-i", path, "-q", fq, "-o", str(tmp_path / "x"), "--gpus", "3"], capture_output=True, text=True)
    assert r.returncode == 1 and "--gpus must be in [1,2]" in r.stderr


@pytest.mark.gpu
def test_cli_two_real_gpus(built_lib, tmp_path):
    """the shipped tool with --gpus 2 on two REAL devices (skipped on a one-GPU box): image on both, every batch split by
    cumulative k-mer count, results spliced back in read order == the oracle; also with --deduplicate"""
    import fulgor_b200 as fg

    if fg.lib().fulgor_gpu_device_count() < 2:
        pytest.skip("needs two CUDA devices")
    reads = ck.gen_reads(30000, 75, 300, seed=63, genomes="synth_200")
    fq = str(tmp_path / "reads.fq")
    write_fastq(fq, reads)
    path = ck.index_path("synth_200.mfur")
    o = ck.Oracle(path)
    for args, algo, thr in ((["--gpus", "2"], 0, 1.0), (["--gpus", "2", "-r", "0.6"], 1, 0.6), (["--gpus", "2", "--deduplicate", "-t", "6"], 0, 1.0)):
        out = str(tmp_path / "out.txt")
        subprocess.check_call([GPU_CLI, "-i", path, "-q", fq, "-o", out, "--batch-reads", "7000"] + args)
        assert ascii_records(out) == csr_records(o.pseudoalign(reads, algo, thr))


def test_cli_flag_errors(built_lib, tmp_path):
    CLI = GPU_CLI
    """flag validation mirrors tools/pseudoalign.cpp:272-321 and needs no GPU"""
    if not os.path.exists(CLI):
        pytest.skip("CLI not built")
    r = subprocess.run([CLI, "-i", "x.fur", "-q", "q.fq", "-o", "o", "-r", "1.5"], capture_output=True, text=True)
    assert r.returncode == 1 and "threshold must be a float in (0.0,1.0]" in r.stderr
    r = subprocess.run([CLI, "-i", "x.txt", "-q", "q.fq", "-o", "o"], capture_output=True, text=True)
    assert r.returncode == 1 and "Wrong index filename supplied." in r.stderr
    r = subprocess.run([CLI, "-i", "x.fur", "-q", "q.fq", "-o", "o", "--format", "xml"], capture_output=True, text=True)
    assert r.returncode == 1 and "Unknown output format" in r.stdout
    r = subprocess.run([CLI, "-q", "q.fq", "-o", "o"], capture_output=True, text=True)
    assert r.returncode == 1


REF_GPU_CLI = os.path.join(ck.ROOT, "oracle", "_ref", "fulgor_ref_gpu")


@pytest.mark.gpu
@pytest.mark.parametrize("index,algo_args", [("salmonella_10.fur", []), ("salmonella_10.mfur", ["-r", "0.8"]), ("synth_200.dfur", []),
                                             ("synth_skew.fur", ["-r", "0.6"])])
def test_reference_tool_with_the_gpu_worker(index, algo_args, tmp_path, built_lib):
    """the drop-in claim as a test: oracle/_ref/fulgor_ref_gpu is the REFERENCE's pseudoalign tool (its parser, FQFeeder, formatters,
    counters, compiled from /root/reference) with pseudoalign_worker replaced by one fulgor_gpu_pseudoalign call per chunk
    (oracle/ref/ref_gpu_cli.cpp = INTEGRATION.md compiled). Its output must equal the unmodified reference binary's, record
    for record, in all three formats (sorted by read id: both write in thread order)."""
    if not (os.path.exists(REF_GPU_CLI) and os.path.exists(ck.REF_CLI)):
        pytest.skip("oracle/_ref/fulgor_ref_gpu not built (make -C oracle ref_gpu; needs /root/reference at build time)")
    genomes = index.split(".")[0]
    reads = ck.gen_reads(20000, 75, 300, seed=51, genomes=genomes)
    bases, off = reads
    seqs = [bases[int(off[i]):int(off[i + 1])].tobytes() for i in range(len(off) - 1)]
    seqs[7] = seqs[7][:40] + b"N" + seqs[7][41:]
    seqs[11] = b"ACGT" * 5
    reads = ck.reads_from_list(seqs)
    fq = str(tmp_path / "reads.fq")
    write_fastq(fq, reads)
    path = ck.index_path(index)
    for fmt, parse in (("ascii", ascii_records), ("binary", binary_records), ("compressed", compressed_records)):
        a, b = str(tmp_path / f"gpu.{fmt}"), str(tmp_path / f"ref.{fmt}")
        subprocess.check_call([REF_GPU_CLI, "pseudoalign", "-i", path, "-q", fq, "-o", a, "-t", "4", "--format", fmt] + algo_args,
                              stdout=subprocess.DEVNULL)
        subprocess.check_call([ck.REF_CLI, "pseudoalign", "-i", path, "-q", fq, "-o", b, "-t", "4", "--format", fmt] + algo_args,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        ra, rb = parse(a), parse(b)
        assert ra == rb
        assert len(ra if fmt != "compressed" else ra[1]) == len(seqs)
