"""Run the UNMODIFIED reference `run_crispresso` (CRISPResso/CRISPRessoCORE.py:1216) over stand-in `needle` / `flash`
executables on PATH -- TEST INFRASTRUCTURE.

The reference shells out to `needle` four times and to `flash` once per run (CORE:1655-1664, 1791-1936) and parses
their output with its own code (parse_needle_output CORE:1707-1786, the FASTQ counters CORE:313-348).  This module
builds a directory of executables with those names and puts it first on PATH, in one of three flavours:

  gpu      crispresso_b200/bin/{needle,flash}: libcrgpu behind the reference's own command lines (needs a B200);
  oracle   the same command-line front ends (crispresso_b200.needle_cli / flash_cli: argument parsing, FASTA
           reading, srspair writer) with the CPU oracle computing the alignments / merges -- what the CPU test
           box can run;
  replay   serves outputs captured from the `gpu` flavour on a B200 (tests/golden/dropin_capture/), keyed by a
           hash of the command line's inputs: the reference then consumes bytes the GPU produced.

With CRGPU_DROPIN_CAPTURE=<dir> every needle / flash invocation also stores its output there under that key, and
with CRGPU_DROPIN_REQUESTS=<dir> its inputs (command line, amplicon FASTA, the FASTA stream on stdin): the requests the
unmodified reference issues are committed under tests/golden/dropin_requests/ (tests/golden/make_dropin_requests.py,
run where /root/reference exists) and answered on a B200 by scripts/make_dropin_capture.py.
"""
import gzip
import hashlib
import json
import os
import stat
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REQUEST_FASTA_DIR = "/tmp/crgpu_dropin_requests"
RUNDATE = "Mon 19 Oct 2026 00:00:00"        # pins the only non-deterministic line of the srspair header


# ---------------------------------------------------------------------------------- back ends
def oracle_align(amplicon, reads, gapopen, gapextend):
    from oracle import needle
    res, ref, mark, qry = needle.align_batch(amplicon, reads, gapopen, gapextend, nthreads=os.cpu_count() or 1)
    return res["ident"], res["score"], ref, mark, qry


def oracle_merge(r1, r2, outdir, o):
    """FLASH restatement + the files flash.flash_merge_files writes."""
    from crispresso_b200.flash import combined_tag
    from oracle import fastq, flash_merge
    a, b = fastq.read_fastq(r1), fastq.read_fastq(r2)
    lens = []
    with gzip.open(os.path.join(outdir, "out.extendedFrags.fastq.gz"), "wt") as fe, \
            gzip.open(os.path.join(outdir, "out.notCombined_1.fastq.gz"), "wt") as f1, \
            gzip.open(os.path.join(outdir, "out.notCombined_2.fastq.gz"), "wt") as f2:
        for (h1, s1, q1), (h2, s2, q2) in zip(a, b):
            m = flash_merge.merge_pair(s1, q1, s2, q2, min_overlap=o["min_overlap"], max_overlap=o["max_overlap"],
                                       max_mismatch_density=o["max_mismatch_density"], allow_outies=o["allow_outies"])
            if m:
                fe.write("@%s\n%s\n+\n%s\n" % (combined_tag(h1, h2), m[0], m[1]))
                lens.append(len(m[0]))
            else:
                f1.write("@%s\n%s\n+\n%s\n" % (h1, s1, q1))
                f2.write("@%s\n%s\n+\n%s\n" % (h2, s2, q2))
    return os.path.join(outdir, "out.extendedFrags.fastq.gz"), lens


# ---------------------------------------------------------------------------------- capture keys
def needle_key(argv, stdin_bytes):
    """Hash of what a needle run depends on: the amplicon FASTA's content, the option tokens, the FASTA stream."""
    h = hashlib.sha256()
    for tok in argv:
        if tok.startswith("-asequence="):
            with open(tok.split("=", 1)[1], "rb") as f:
                h.update(b"A" + f.read())
        elif tok.startswith(("-bsequence=", "-outfile=")):
            continue
        else:
            h.update(b"O" + tok.encode())
    h.update(b"S" + stdin_bytes)
    return "needle_" + h.hexdigest()[:24]


def flash_key(argv):
    from crispresso_b200.flash_cli import parse_command_line
    files, o = parse_command_line(argv)
    h = hashlib.sha256()
    for fn in files:
        with gzip.open(fn, "rb") if fn.endswith(".gz") else open(fn, "rb") as f:
            h.update(b"F" + f.read())
    h.update(repr(sorted((k, v) for k, v in o.items() if k not in ("outdir",))).encode())
    return "flash_" + h.hexdigest()[:24], o


def save_needle_request(directory, argv, stdin_bytes):
    """One needle invocation of the reference, self-contained: tokens with the amplicon FASTA inlined."""
    os.makedirs(directory, exist_ok=True)
    afasta, toks = "", []
    for tok in argv:
        if tok.startswith("-asequence="):
            with open(tok.split("=", 1)[1], "rt") as f:
                afasta = f.read()
            toks.append("-asequence=@AFASTA@")
        else:
            toks.append(tok)
    key = needle_key(argv, stdin_bytes)
    with gzip.GzipFile(os.path.join(directory, key + ".req.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps({"argv": toks, "afasta": afasta, "stdin": stdin_bytes.decode("latin-1")}).encode())
    with open(os.path.join(directory, "order.log"), "at") as f:
        f.write(key + "\n")
    return key


def request_reads(path):
    """(names, sequences) of the FASTA stream of a recorded needle request, and its amplicon / option tokens."""
    import io

    from crispresso_b200 import needle_cli
    with gzip.open(path, "rb") as f:
        d = json.loads(f.read().decode())
    names, seqs = needle_cli.read_fasta(io.BytesIO(d["stdin"].encode("latin-1")))
    _an, aseq = needle_cli.read_fasta(io.StringIO(d["afasta"]))
    return names, seqs, aseq[0], d["argv"]


def load_needle_request(path, workdir):
    """-> (key, argv with a real amplicon FASTA path, stdin bytes)"""
    with gzip.open(path, "rb") as f:
        d = json.loads(f.read().decode())
    key = os.path.basename(path)[:-len(".req.json.gz")]
    # a fixed place: the path is echoed in the srspair header ("#    -asequence=..."), and the answers must not depend on it
    os.makedirs(REQUEST_FASTA_DIR, exist_ok=True)
    fa = os.path.join(REQUEST_FASTA_DIR, key + "_a.fa")
    with open(fa, "wt") as f:
        f.write(d["afasta"])
    argv = [("-asequence=" + fa) if t == "-asequence=@AFASTA@" else t for t in d["argv"]]
    return key, argv, d["stdin"].encode("latin-1")


# ---------------------------------------------------------------------------------- executables
def needle_main(flavour, argv):
    import io

    from crispresso_b200 import needle_cli
    os.environ.setdefault("CRGPU_NEEDLE_RUNDATE", RUNDATE)
    data = sys.stdin.buffer.read()
    cap = os.environ.get("CRGPU_DROPIN_CAPTURE")
    if os.environ.get("CRGPU_DROPIN_REQUESTS"):
        try:
            needle_cli.parse_command_line(argv)
            save_needle_request(os.environ["CRGPU_DROPIN_REQUESTS"], argv, data)
        except needle_cli.UsageError:
            pass                                        # (the malformed repair-RC command: nothing to answer)
    if flavour == "replay":
        try:
            key = needle_key(argv, data)
        except Exception:
            key = None
        try:
            needle_cli.parse_command_line(argv)
        except needle_cli.UsageError as e:              # the malformed repair-RC command: fails like the real thing
            sys.stderr.write("Died: %s\n" % e)
            return 1
        path = os.path.join(os.environ["CRGPU_DROPIN_REPLAY"], key + ".txt.gz")
        if not os.path.exists(path):
            sys.stderr.write("replay: no capture %s\n" % path)
            return 1
        with gzip.open(path, "rb") as f:
            sys.stdout.buffer.write(f.read())
        return 0
    align = oracle_align if flavour == "oracle" else needle_cli.gpu_align
    real_stdout = sys.stdout
    buf = io.StringIO()
    sys.stdout = buf
    try:
        rc = needle_cli.main(argv, align=align, stdin=io.BytesIO(data))
    finally:
        sys.stdout = real_stdout
    text = buf.getvalue().encode()
    real_stdout.buffer.write(text)
    real_stdout.flush()
    if cap and rc == 0:
        os.makedirs(cap, exist_ok=True)
        with gzip.GzipFile(os.path.join(cap, needle_key(argv, data) + ".txt.gz"), "wb", mtime=0) as f:
            f.write(text)
    return rc


def flash_main(flavour, argv):
    import shutil

    from crispresso_b200 import flash_cli
    cap = os.environ.get("CRGPU_DROPIN_CAPTURE")
    if os.environ.get("CRGPU_DROPIN_REQUESTS"):
        key, o = flash_key(argv)
        files, _o = flash_cli.parse_command_line(argv)
        os.makedirs(os.environ["CRGPU_DROPIN_REQUESTS"], exist_ok=True)
        with open(os.path.join(os.environ["CRGPU_DROPIN_REQUESTS"], key + ".req.json"), "wt") as f:
            toks = [("@R1@" if t == files[0] else "@R2@" if t == files[1] else "@OUTDIR@" if t == o["outdir"] else t) for t in argv]
            json.dump({"argv": toks, "r1": os.path.basename(files[0]), "r2": os.path.basename(files[1])}, f)
        with open(os.path.join(os.environ["CRGPU_DROPIN_REQUESTS"], "order.log"), "at") as f:
            f.write(key + "\n")
    if flavour == "replay":
        key, o = flash_key(argv)
        src = os.path.join(os.environ["CRGPU_DROPIN_REPLAY"], key + ".extendedFrags.fastq.gz")
        if not os.path.exists(src):
            sys.stderr.write("replay: no capture %s\n" % src)
            return 1

        def merge(_r1, _r2, outdir, _o):
            dst = os.path.join(outdir, "out.extendedFrags.fastq.gz")
            shutil.copyfile(src, dst)
            from oracle import fastq
            for part in ("1", "2"):
                with gzip.open(os.path.join(outdir, "out.notCombined_%s.fastq.gz" % part), "wt"):
                    pass
            return dst, [len(s) for _h, s, _q in fastq.read_fastq(dst)]
        return flash_cli.main(argv, merge=merge)
    merge = oracle_merge if flavour == "oracle" else flash_cli.gpu_merge
    rc = flash_cli.main(argv, merge=merge)
    if cap and rc == 0:
        key, o = flash_key(argv)
        os.makedirs(cap, exist_ok=True)
        # re-compress with a fixed mtime so that the capture is reproducible
        with gzip.open(os.path.join(o["outdir"], "%s.extendedFrags.fastq%s" % (o["prefix"], ".gz" if o["gz"] else "")), "rb") as f:
            data = f.read()
        with gzip.GzipFile(os.path.join(cap, key + ".extendedFrags.fastq.gz"), "wb", mtime=0) as f:
            f.write(data)
    return rc


def make_bin_dir(path, flavour):
    """Directory with `needle`, `flash`, `java` executables of the given flavour."""
    assert flavour in ("gpu", "oracle", "replay")
    os.makedirs(path, exist_ok=True)
    for exe in ("needle", "flash"):
        p = os.path.join(path, exe)
        with open(p, "w") as f:
            f.write("#!%s\nimport sys\nsys.path[:0] = [%r, %r]\nfrom dropin import harness\n"
                    "sys.exit(harness.%s_main(%r, sys.argv[1:]))\n" % (sys.executable, ROOT, os.path.join(ROOT, "tests"), exe, flavour))
        os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC | stat.S_IXGRP | stat.S_IXOTH)
    p = os.path.join(path, "java")
    with open(p, "w") as f:
        f.write("#!/bin/sh\nexit 0\n")
    os.chmod(p, os.stat(p).st_mode | stat.S_IEXEC)
    return path


# ---------------------------------------------------------------------------------- the reference itself
def run_reference(bin_dir, out_dir, fastq_r1, fastq_r2="", amplicon_seq="", guide_seq="", n_processes=1,
                  keep_intermediate=False, extra=None, cwd=None):
    """args as tests/crispresso_tests.py builds them (parse_args over a hijacked sys.argv, then attribute
    assignments) -> the 14-tuple of run_crispresso (CORE:3977-3992)."""
    import ref_shim
    core = ref_shim.load_core(rich_plots=True)
    old_path, old_argv, old_cwd = os.environ.get("PATH", ""), sys.argv, os.getcwd()
    os.environ["PATH"] = bin_dir + os.pathsep + old_path
    sys.argv = ["CRISPResso", "-r1", fastq_r1, "--amplicon_seq", amplicon_seq or "ACGT"]
    try:
        if cwd:
            os.chdir(cwd)
        args = core.parse_args(sys.argv[1:])
        args.fastq_r1, args.fastq_r2 = fastq_r1, fastq_r2
        args.amplicon_seq, args.guide_seq = amplicon_seq, guide_seq
        args.n_processes, args.keep_intermediate = n_processes, keep_intermediate
        args.output_folder = out_dir
        args.trim_sequences = False
        for k, v in (extra or {}).items():
            setattr(args, k, v)
        return core.run_crispresso(args)
    finally:
        os.environ["PATH"], sys.argv = old_path, old_argv
        os.chdir(old_cwd)
