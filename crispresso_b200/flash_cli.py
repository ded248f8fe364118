"""`flash` command line backed by libcrgpu -- the drop-in for the FLASH subprocess.

CRISPResso shells out to
    flash R1 R2 --allow-outies --max-overlap M --min-overlap m -f <amplicon len> -r <read len> -s <sd> -z -d DIR
(CRISPResso/CRISPRessoCORE.py:1655-1664) and goes on with DIR/out.extendedFrags.fastq.gz (CORE:1677).  With
`crispresso_b200/bin` first on PATH the UNMODIFIED reference merges its read pairs with crgpu_flash_merge
(FLASH 1.2.11 semantics: SURVEY.md App. D) and finds the files FLASH would have
written: out.extendedFrags / out.notCombined_{1,2} (.fastq or .fastq.gz), out.hist, out.histogram.
-f / -r / -s only steer FLASH's default for --max-overlap, which CRISPResso always passes explicitly.
No CPU fallback: without libcrgpu.so / a B200 the program exits with status 1.
"""
import collections
import gzip
import os
import sys


class UsageError(Exception):
    pass


_VALUE = {"-m": "min_overlap", "--min-overlap": "min_overlap", "-M": "max_overlap", "--max-overlap": "max_overlap",
          "-x": "max_mismatch_density", "--max-mismatch-density": "max_mismatch_density", "-d": "outdir",
          "--output-directory": "outdir", "-o": "prefix", "--output-prefix": "prefix", "-f": None, "--fragment-len": None,
          "-r": None, "--read-len": None, "-s": None, "--fragment-len-stddev": None, "-t": None, "--threads": None,
          "-p": None, "--phred-offset": None}
_FLAG = {"-O": "allow_outies", "--allow-outies": "allow_outies", "-z": "gz", "--compress": "gz", "-q": None, "--quiet": None}


def parse_command_line(argv):
    o = dict(min_overlap=10, max_overlap=65, max_mismatch_density=0.25, outdir=".", prefix="out", allow_outies=False, gz=False)
    files = []
    i = 0
    while i < len(argv):
        tok = argv[i]
        key, val = tok, None
        if tok.startswith("--") and "=" in tok:
            key, val = tok.split("=", 1)
        if key in _FLAG:
            if _FLAG[key]:
                o[_FLAG[key]] = True
        elif key in _VALUE:
            if val is None:
                if i + 1 >= len(argv):
                    raise UsageError("flash: option %s needs a value" % key)
                val = argv[i + 1]
                i += 1
            if _VALUE[key]:
                o[_VALUE[key]] = val
        elif tok.startswith("-") and tok != "-":
            raise UsageError("flash: option %s is not supported by the GPU merger" % tok)
        else:
            files.append(tok)
        i += 1
    if len(files) != 2:
        raise UsageError("flash: expected two FASTQ files, got %d" % len(files))
    try:
        o["min_overlap"], o["max_overlap"] = int(o["min_overlap"]), int(o["max_overlap"])
        o["max_mismatch_density"] = float(o["max_mismatch_density"])
    except ValueError:
        raise UsageError("flash: overlap bounds must be numbers")
    return files, o


def gpu_merge(r1, r2, outdir, o, _state={}):
    from . import Context
    from .flash import flash_merge_files
    if "ctx" not in _state:
        _state["ctx"] = Context(0)
    ext, _n1, _n2, res = flash_merge_files(_state["ctx"], r1, r2, outdir, min_overlap=o["min_overlap"],
                                           max_overlap=o["max_overlap"], allow_outies=o["allow_outies"],
                                           max_mismatch_density=o["max_mismatch_density"])
    return ext, [int(x) for x in (res.offsets[1:] - res.offsets[:-1])]


def main(argv=None, merge=gpu_merge):
    argv = sys.argv[1:] if argv is None else argv
    try:
        files, o = parse_command_line(argv)
        os.makedirs(o["outdir"], exist_ok=True)
        ext, merged_lens = merge(files[0], files[1], o["outdir"], o)
    except UsageError as e:
        sys.stderr.write("%s\n" % e)
        return 1
    except Exception as e:
        sys.stderr.write("flash: %s\n" % e)
        return 1
    base = os.path.join(o["outdir"], o["prefix"])
    # the mergers write out.*.fastq.gz; honour a different prefix / no -z by renaming / inflating
    for part in ("extendedFrags", "notCombined_1", "notCombined_2"):
        src = os.path.join(o["outdir"], "out.%s.fastq.gz" % part)
        dst = "%s.%s.fastq%s" % (base, part, ".gz" if o["gz"] else "")
        if src == dst or not os.path.exists(src):
            continue
        if o["gz"]:
            os.replace(src, dst)
        else:
            with gzip.open(src, "rb") as f, open(dst, "wb") as g:
                g.write(f.read())
            os.remove(src)
    hist = collections.Counter(merged_lens)
    with open(base + ".hist", "wt") as f:
        for ln in sorted(hist):
            f.write("%d\t%d\n" % (ln, hist[ln]))
    top = max(hist.values()) if hist else 1
    with open(base + ".histogram", "wt") as f:
        for ln in sorted(hist):
            f.write("%d\t%s\n" % (ln, "*" * max(1, int(round(72.0 * hist[ln] / top)))))
    sys.stdout.write("[FLASH] Combined pairs: %d (libcrgpu)\n" % len(merged_lens))
    return 0


if __name__ == "__main__":
    sys.exit(main())
