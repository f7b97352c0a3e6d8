"""CPU restatement of the text loop of EdgeCalculator::construct_edges (src/EdgeCalculator.cpp:581-645)
and of the Overlap constructor (src/Overlap.h:39-73) -- TEST INFRASTRUCTURE ONLY (tests/, smoke).

Pure Python, small inputs.  Pinned against the unmodified reference through tests/golden/ingest_*.npz
(what the real construct_edges() wrote to nonedge_overlaps.txt for a file of irregular but accepted
lines, with thresholds under which every surviving line is printed back)."""
from __future__ import annotations

from typing import Dict, List, Tuple

ULONG_MAX = 2 ** 64 - 1
_SPACE = b" \t\n\v\f\r"

SCORE, NONEDGE, DROPPED, SKIPPED, ERROR, UNKNOWN_ID = 1, 2, 3, 4, 5, 6


def c_strtoul0(b: bytes) -> int:
    """glibc strtoul(s, NULL, 0), the body of str_to_read_id (src/Types.h:99-102)."""
    k = 0
    while k < len(b) and b[k] in _SPACE:
        k += 1
    neg = False
    if k < len(b) and b[k] in b"+-":
        neg = b[k] == ord("-")
        k += 1
    base = 10
    if k < len(b) and b[k] == ord("0"):
        if k + 2 < len(b) and b[k + 1] in b"xX" and chr(b[k + 2]) in "0123456789abcdefABCDEF":
            base = 16
            k += 2
        else:
            base = 8
    v = 0
    ovf = False
    while k < len(b):
        c = chr(b[k])
        if c.isdigit() and c.isascii():
            d = ord(c) - 48
        elif "a" <= c <= "z":
            d = ord(c) - 87
        elif "A" <= c <= "Z":
            d = ord(c) - 55
        else:
            break
        if d >= base:
            break
        v = v * base + d
        if v > ULONG_MAX:
            ovf = True
        k += 1
    if ovf:
        return ULONG_MAX
    return (-v) % 2 ** 64 if neg else v


def c_atoi_u32(b: bytes) -> int:
    """(unsigned int)atoi(s), atoi = (int)strtol(s, NULL, 10) (src/Overlap.h:42-50)."""
    k = 0
    while k < len(b) and b[k] in _SPACE:
        k += 1
    neg = False
    if k < len(b) and b[k] in b"+-":
        neg = b[k] == ord("-")
        k += 1
    v = 0
    while k < len(b) and 48 <= b[k] <= 57:
        v = v * 10 + b[k] - 48
        k += 1
    r = -v if neg else v
    r = max(-(2 ** 63), min(2 ** 63 - 1, r))      # strtol saturates
    return r % 2 ** 32                             # (int) then (unsigned int)


def _as_int(u: int) -> int:
    return u - 2 ** 32 if u >= 2 ** 31 else u


def char_field(b: bytes, is_type: bool) -> int:
    """check_ori / check_ord strip ' ' (src/Overlap.h:115-118,127-130), check_type strips '\\n', '\\t', ' '
    (:152-156) when the length is not 1; the result must be one character (assert).  0 = not one."""
    if len(b) == 1:
        return b[0]
    strip = b" \n\t" if is_type else b" "
    rest = bytes(c for c in b if c not in strip)
    return rest[0] if len(rest) == 1 else 0


def split_lines(text: bytes) -> List[bytes]:
    """std::getline over the file: an unterminated last line counts, an empty remainder does not."""
    lines = text.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    return lines


def ingest(text: bytes, id_to_index: Dict[int, int], min_overlap_len: int, min_overlap_perc: int = 0, relax_PE_edges: bool = False,
           allow_spaces: bool = False, max_overlaps: int = ULONG_MAX):
    """Returns (status per line, scored [(line, rec, idx1, idx2)], filtered [(line, rec)]); rec is the tuple
    (id1, id2, pos1, pos2, ord, ori1, ori2, perc1, perc2, len1, len2, type1, type2) with characters as ints."""
    status, scored, filtered = [], [], []
    for i, line in enumerate(split_lines(text)):
        if i >= max_overlaps:                                                   # :581
            break
        line = line.strip(b"\t ")                                               # :584
        if allow_spaces:                                                        # :585-587, token_compress_on
            f, cur, k = [], b"", 0
            while k < len(line):
                if line[k] in b"\t ":
                    f.append(cur)
                    cur = b""
                    while k + 1 < len(line) and line[k + 1] in b"\t ":
                        k += 1
                else:
                    cur += line[k:k + 1]
                k += 1
            f.append(cur)
        else:
            f = line.split(b"\t") if line else []                               # :589-593
        if len(f) != 13:                                                        # :598-603
            status.append(SKIPPED)
            continue
        id1, id2 = c_strtoul0(f[0]), c_strtoul0(f[1])
        pos1, pos2 = c_atoi_u32(f[2]), c_atoi_u32(f[3])
        perc1, perc2, len1, len2 = c_atoi_u32(f[7]), c_atoi_u32(f[8]), c_atoi_u32(f[9]), c_atoi_u32(f[10])
        if f[3] == b"-":                                                        # src/Overlap.h:55-59
            pos2 = perc2 = len2 = 0
        od, o1, o2 = char_field(f[4], False), char_field(f[5], False), char_field(f[6], False)
        t1, t2 = char_field(f[11], True), char_field(f[12], True)
        bad = _as_int(pos1) < 0 or _as_int(pos2) < 0                            # check_pos :107-112
        bad |= o1 not in b"+-" or o2 not in b"+-" or o1 == 0 or o2 == 0        # check_ori :125-134
        bad |= not (0 <= _as_int(perc1) <= 100) or not (0 <= _as_int(perc2) <= 100)   # check_perc :136-142
        bad |= _as_int(len1) < 0 or _as_int(len2) < 0                           # check_len :144-149
        bad |= t1 not in b"sp" or t2 not in b"sp" or t1 == 0 or t2 == 0        # check_type :151-163
        if not bad:                                                             # check_ord :114-123
            if t1 == ord("s") or t2 == ord("s"):
                bad = od != ord("-")
            else:
                bad = od not in (ord("1"), ord("2"))
        if bad:
            status.append(ERROR)
            continue
        rec = (id1, id2, pos1, pos2, od, o1, o2, perc1, perc2, len1, len2, t1, t2)
        if id1 == id2:                                                          # :605-607
            status.append(DROPPED)
            continue
        perc = int(0.5 * ((perc1 + perc2) % 2 ** 32)) if perc2 > 0 else perc1   # Overlap::get_perc :203-210
        any_p = t1 == ord("p") or t2 == ord("p")
        if len1 >= min_overlap_len and not any_p:                               # :612-617
            band = True
        elif len1 >= 0.5 * min_overlap_len and len2 >= 0.5 * min_overlap_len and any_p:   # :618-624
            band = True
        else:
            band = relax_PE_edges and (len1 + len2) % 2 ** 32 >= min_overlap_len and any_p   # :626-632
        if not band:
            status.append(NONEDGE)                                              # :633-635
            filtered.append((i, rec))
            continue
        if perc < min_overlap_perc:
            status.append(DROPPED)
            continue
        if id1 not in id_to_index or id2 not in id_to_index:                    # map::at throws, :170-171
            status.append(UNKNOWN_ID)
            continue
        status.append(SCORE)
        scored.append((i, rec, id_to_index[id1], id_to_index[id2]))
    return status, scored, filtered


def rec_line(rec: Tuple) -> str:
    """Overlap::get_overlap_line (src/Overlap.h:234-237) without the newline."""
    return "%d\t%d\t%d\t%d\t%s\t%s\t%s\t%d\t%d\t%d\t%d\t%s\t%s" % (rec[0], rec[1], rec[2], rec[3], chr(rec[4]), chr(rec[5]), chr(rec[6]),
                                                                   rec[7], rec[8], rec[9], rec[10], chr(rec[11]), chr(rec[12]))
