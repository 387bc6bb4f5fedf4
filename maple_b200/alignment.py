"""MAPLE-format alignments in, tip genome lists out: the producers of the search's inputs.

`read_maple_alignment` mirrors readConciseAlignment (MAPLEv0.7.5.4.py:3498-3553): a reference genome record followed by one
record per sample holding only its differences, one per line: `char<TAB>pos[<TAB>length]` (`n`/`-` runs carry a length;
1-based positions); plain text or .gz.  `tip_genome_list` mirrors probVectTerminalNode (:3882-3962) for a sample that is not
yet in a tree (node=None): R runs between differences, N runs, substitutions, IUPAC ambiguity codes as O entries (with the
reference's error-model adjustment of the ambiguity vector when usingErrorRate).  Host code, cheap, no GPU involved; the
lists go through genome_list.pack_lists to the device.
"""
from __future__ import annotations

import gzip
from typing import Dict, List, Optional, Sequence, Tuple

ALLELES_LOW = {"a": 0, "c": 1, "g": 2, "t": 3}
# :3666 -- unnormalised 0/1 vectors, as the reference stores them
AMBIGUITIES = {"y": [0.0, 1.0, 0.0, 1.0], "r": [1.0, 0.0, 1.0, 0.0], "w": [1.0, 0.0, 0.0, 1.0], "s": [0.0, 1.0, 1.0, 0.0],
               "k": [0.0, 0.0, 1.0, 1.0], "m": [1.0, 1.0, 0.0, 0.0], "d": [1.0, 0.0, 1.0, 1.0], "v": [1.0, 1.0, 1.0, 0.0],
               "h": [1.0, 1.0, 0.0, 1.0], "b": [0.0, 1.0, 1.0, 1.0]}

Diff = Tuple  # (char, pos) or (char, pos, length)


class AlignmentError(ValueError):
    """The reference prints a message and raises Exception("exit") in these cases (:3527-3539)."""


def _records(text: str):
    """Split a MAPLE/FASTA-like text into (header, body lines) records; reading stops at the first empty line that
    follows a record, as the reference's reader does (:3515, :3520)."""
    header, body = None, []
    for raw in text.split("\n"):
        if raw.startswith(">"):
            if header is not None:
                yield header, body
            header, body = raw[1:].replace(">", ""), []
        elif raw == "":
            if header is not None:
                break
        elif header is not None:
            body.append(raw)
    if header is not None:
        yield header, body


def _parse_diff(line: str, ref: str, where: str) -> Diff:
    cols = line.split()
    if len(cols) < 2:
        raise AlignmentError("%s: line with only one column: %r (is the reference included at the top of the alignment?)" % (where, line))
    ch, pos = cols[0].lower(), int(cols[1])
    if ch not in ("n", "-") and ref[pos - 1] == ch:
        raise AlignmentError("mutation into the reference nucleotide at position %d (%s): wrong reference?" % (pos, ch))
    return (ch, pos, int(cols[2])) if len(cols) > 2 else (ch, pos)


def read_maple_alignment(path: str, reference: Optional[str] = None) -> Tuple[str, Dict[str, List[Diff]]]:
    """Returns (reference genome in lower case, {sample name: [(char, pos[, length]), ...]}).  With `reference` given the file
    is expected to hold samples only (the reference's --reference option, extractReference=False).  Same accepted inputs and
    the same three rejections as readConciseAlignment (:3527-3539): a one-column line, a substitution into the reference
    base, an entry that starts inside the previous one."""
    with (gzip.open if path.endswith(".gz") else open)(path, "rt") as f:
        records = _records(f.read())
    if reference is None:
        first = next(records, None)
        ref = "".join(first[1]).lower() if first else ""
    else:
        ref = reference.lower()
    data: Dict[str, List[Diff]] = {}
    for number, (name, lines) in enumerate(records, 1):
        seq: List[Diff] = []
        covered = 0  # last position the previous entry covers
        for line in lines:
            entry = _parse_diff(line, ref, path)
            if entry[1] <= covered:
                raise AlignmentError("sample %d (%s): entry %r overlaps the previous one %r" % (number, name, line.strip(), seq[-1]))
            covered = entry[1] + (entry[2] - 1 if len(entry) == 3 else 0)
            seq.append(entry)
        data[name] = seq
    return ref, data


def tip_genome_list(diffs: Optional[Sequence[Diff]], refIdx: Sequence[int], lRef: int, usingErrorRate: bool = False,
                    errorRate: float = 0.0, errorRates: Optional[Sequence[float]] = None, onlyNambiguities: bool = False,
                    numMinSeqs: int = 0) -> list:
    """probVectTerminalNode(diffs, None, None): the lower genome list of a new tip, relative to the reference genome.
    numMinSeqs > 0: the tip stands for several identical samples, which under the error model gives its ambiguity vectors
    no error term (updateProbVectTerminalNode, :3980-4003)."""
    if diffs is None:
        return [(5, lRef)]
    pos = 1
    out: list = []
    for m in diffs:
        cur = m[1]
        if cur > pos:  # identical to the reference up to here
            out.append((4, cur - 1))
            pos = cur
        ch = m[0]
        if ch == "n" or ch == "-":
            length = m[2] if len(m) > 2 else 1
            entry = (5, cur + length - 1)
            pos = cur + length
        elif ch in ALLELES_LOW:
            if ALLELES_LOW[ch] == refIdx[cur - 1]:  # the reference warns and stores an R entry (:3904-3908)
                entry = (4, cur)
            else:
                entry = (ALLELES_LOW[ch], int(refIdx[cur - 1]))
            pos = cur + 1
        else:
            if onlyNambiguities:
                entry = (5, cur)
            elif usingErrorRate:  # numMinSeqs == 0 for a new tip
                # What tips look like under the error model: [0.5 - eps/3, eps/3, ..] for two states, [1/3 - eps/9, eps/3] for
                # three (updateProbVectTerminalNode, :3980-4003).  probVectTerminalNode itself starts from the 0/1 table
                # (:3923-3938), but that table is shared by reference with the tips already placed and is rewritten in place by
                # the first updateProbVectTerminalNode, so from then on -- in every SPR round -- the reference produces exactly
                # the values below (pinned by tests/golden/*err*).
                base = AMBIGUITIES[ch]
                n_set = sum(bool(x) for x in base)
                eps = float(errorRates[cur - 1]) if errorRates is not None else float(errorRate)
                if numMinSeqs and n_set in (2, 3):
                    vec = [0.0 if x == 0 else (0.5 if n_set == 2 else 1.0 / 3) for x in base]
                elif n_set == 2:
                    vec = [eps * 0.33333 if x == 0 else 0.5 - eps * 0.33333 for x in base]
                elif n_set == 3:
                    vec = [eps * 0.33333 if x == 0 else (1.0 / 3) - eps / 9 for x in base]
                else:
                    vec = list(base)
                entry = (6, int(refIdx[cur - 1]), vec)
            else:
                entry = (6, int(refIdx[cur - 1]), list(AMBIGUITIES[ch]))
            pos = cur + 1
        out.append(entry)
    if pos <= lRef:
        out.append((4, lRef))
    return out
