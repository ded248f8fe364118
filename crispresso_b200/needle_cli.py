"""`needle` command line backed by libcrgpu -- the drop-in for the EMBOSS needle subprocess.

CRISPResso shells out to
    ... | sed 's/:/_/g' | needle -asequence=<amplicon.fa> -bsequence=/dev/stdin -outfile=/dev/stdout
          -gapopen=10 -gapextend=0.5 -awidth3=5000 2>> log | gzip > needle_output_<id>.txt.gz
four times per run (CRISPResso/CRISPRessoCORE.py:1791-1806, 1812-1824, 1911-1921, 1924-1936) and re-reads the
srspair text with parse_needle_output (CORE:1707-1786).  Putting `crispresso_b200/bin` first on PATH makes the
UNMODIFIED reference run its alignments on the GPU: this program reads the FASTA stream, aligns every
sequence with crgpu_align (needle semantics, bit for bit) and writes the srspair text needle would write.

Behaviour the reference relies on and that is kept:
  * sequence names are the first whitespace-delimited token of the FASTA header;
  * gap characters in the input are dropped (EMBOSS reads `nucleotide` sequences without gaps): the
    reverse-complement rescue feeds aligned reads with their '-' left in (CORE:1846, 1867);
  * a command line with a token that is not a qualifier fails without writing a record -- the repair-RC
    command of CORE:1924-1936 carries the literal text `args.needle_options_string`, and the reference's
    NaN `score_repaired` for `_RC` rows depends on that failure.
There is no CPU fallback: without libcrgpu.so / a B200 the program exits with status 1.
"""
import sys

from . import srspair

CHUNK = 1 << 17                     # reads per crgpu_align call

_KNOWN = {"asequence", "bsequence", "outfile", "gapopen", "gapextend", "awidth3", "auto", "aformat3", "aformat"}


class UsageError(Exception):
    pass


def parse_command_line(argv):
    """EMBOSS qualifiers as CRISPResso writes them: -name=value (or -name value).  Anything else is an error."""
    opts = {}
    i = 0
    while i < len(argv):
        tok = argv[i]
        if not tok.startswith("-") or tok == "-":
            raise UsageError("needle: unexpected parameter %r" % tok)
        body = tok.lstrip("-")
        if "=" in body:
            key, val = body.split("=", 1)
        elif body.lower() == "auto":
            key, val = body, "Y"
        else:
            if i + 1 >= len(argv):
                raise UsageError("needle: qualifier -%s needs a value" % body)
            key, val = body, argv[i + 1]
            i += 1
        key = key.lower()
        if key not in _KNOWN:
            raise UsageError("needle: qualifier -%s is not supported by the GPU aligner" % key)
        opts[key] = val
        i += 1
    for need in ("asequence", "bsequence", "outfile"):
        if need not in opts:
            raise UsageError("needle: -%s is required" % need)
    if opts.get("aformat3", opts.get("aformat", "srspair")).lower() not in ("srspair", "pair"):
        raise UsageError("needle: only the srspair output format is supported")
    try:
        gapopen, gapextend = float(opts.get("gapopen", 10.0)), float(opts.get("gapextend", 0.5))
    except ValueError:
        raise UsageError("needle: gap penalties must be numbers")
    return opts, gapopen, gapextend


def read_fasta(stream):
    """-> (names, sequences): name = first token of the header, gaps and white space dropped."""
    names, seqs, cur = [], [], None
    for line in stream:
        if isinstance(line, bytes):
            line = line.decode("utf-8", "replace")
        line = line.rstrip("\r\n")
        if line.startswith(">"):
            if cur is not None:
                seqs.append("".join(cur))
            tok = line[1:].split()
            names.append(tok[0] if tok else "")
            cur = []
        elif cur is not None:
            cur.append(line.replace("-", "").replace(" ", "").replace(".", ""))
    if cur is not None:
        seqs.append("".join(cur))
    return names, seqs


def gpu_align(amplicon, reads, gapopen, gapextend, _state={}):
    """-> (ident[], score[], ref_rows, markup_rows, read_rows) through crgpu_align."""
    from . import Context
    from .aligner import needle_align
    if "ctx" not in _state:
        _state["ctx"] = Context(0)
    recs, ref, mark, qry = needle_align(_state["ctx"], amplicon, reads, gapopen, gapextend)
    return recs["ident"], recs["score"], ref, mark, qry


def main(argv=None, align=gpu_align, stdin=None):
    argv = sys.argv[1:] if argv is None else argv
    try:
        opts, gapopen, gapextend = parse_command_line(argv)
        with open(opts["asequence"], "rt") as f:
            anames, aseqs = read_fasta(f)
        if len(aseqs) != 1 or not aseqs[0]:
            raise UsageError("needle: -asequence must hold exactly one sequence")
        if opts["bsequence"] in ("/dev/stdin", "stdin", "-"):
            bnames, bseqs = read_fasta(stdin or sys.stdin.buffer)
        else:
            with open(opts["bsequence"], "rb") as f:
                bnames, bseqs = read_fasta(f)
    except (UsageError, OSError) as e:
        sys.stderr.write("Died: %s\n" % e)
        return 1
    out = sys.stdout if opts["outfile"] in ("/dev/stdout", "stdout", "-") else open(opts["outfile"], "wt")
    try:
        out.write(srspair.file_header(opts["asequence"], opts["bsequence"], opts["outfile"], gapopen, gapextend,
                                      opts.get("awidth3")))
        for lo in range(0, len(bseqs), CHUNK):
            chunk = bseqs[lo:lo + CHUNK]
            try:
                ident, score, ref, mark, qry = align(aseqs[0], chunk, gapopen, gapextend)
            except Exception as e:                      # CRGPU_E_ALIGN etc.: as a needle that dies mid-run
                sys.stderr.write("Died: %s\n" % e)
                return 1
            for i in range(len(chunk)):
                out.write(srspair.format_record(anames[0], bnames[lo + i], ref[i], mark[i], qry[i], int(ident[i]),
                                                float(score[i]), gapopen, gapextend))
        out.write(srspair.FILE_TRAILER)
    finally:
        if out is not sys.stdout:
            out.close()
        else:
            out.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
