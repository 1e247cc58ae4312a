"""Shared helpers of the GPU parity tests: run engine and oracle on the same batch and diff."""
import numpy as np

import _oracle as orc
from biogarden_b200 import native, score as score_mod

SCORERS = {"blosum62": score_mod.blosum62, "pam250": score_mod.pam250, "unit": score_mod.unit}


def table256_for(scorer_name, match_mismatch=None):
    """Oracle-side 256x256 table for non-shipped scorers."""
    if match_mismatch is None:
        return None
    mt, mm = match_mismatch
    t = np.full((256, 256), mm, np.int32)
    np.fill_diagonal(t, mt)
    return t


def engine_align(aligner, batch, mode, scorer, a, b, match_mismatch=None, score_only=False):
    sc = score_mod.match_mismatch(*match_mismatch) if match_mismatch else SCORERS[scorer]
    return aligner.align_batch_raw(batch, mode, sc, a, b, score_only=score_only)


def oracle_align(batch, mode, scorer, a, b, match_mismatch=None, lean=True, threads=None, want_strings=True,
                 fresh=True):
    if match_mismatch:
        return orc.align_batch(mode, batch.residues, batch.seq_off, "table", a, b,
                               table=table256_for(scorer, match_mismatch), threads=threads or orc.hw_threads(),
                               lean=lean, want_strings=want_strings, fresh=fresh)
    return orc.align_batch(mode, batch.residues, batch.seq_off, scorer, a, b, threads=threads or orc.hw_threads(),
                           lean=lean, want_strings=want_strings, fresh=fresh)


def fnv_pairs(res: native.Result):
    """FNV-1a-64 over a_align then b_align per pair (same digest the oracle batch driver emits)."""
    out = np.zeros(res.n_pairs, np.uint64)
    arena = res.arena
    for p in range(res.n_pairs):
        h = 14695981039346656037
        for c in arena[int(res.off[2 * p]):int(res.off[2 * p + 2])].tolist():
            h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        out[p] = h
    return out


def diff(batch, eng: native.Result, ora: dict, label="", max_report=5, check_strings=True):
    """Returns a list of human-readable mismatch lines (empty = parity)."""
    msgs = []
    n_bad = 0
    undefined = np.isin(ora["status"], orc.UNDEFINED)
    for p in range(batch.n_pairs):
        o_st = int(ora["status"][p]); e_st = int(eng.status[p])
        bad = None
        if undefined[p]:
            if e_st != native.ST_REF_UNDEFINED:
                bad = "status: oracle says reference-undefined (%d), engine says ok" % o_st
        elif o_st != orc.OK:
            bad = "oracle status %d on a pair the engine accepted" % o_st
        else:
            if e_st != native.ST_OK:
                bad = "status: engine flags REF_UNDEFINED, oracle ok"
            elif int(eng.score[p]) != int(ora["score"][p]):
                bad = "score %d != oracle %d" % (eng.score[p], ora["score"][p])
            elif check_strings:
                ea, eb = eng.strings(p)
                oa, ob = orc.batch_strings(ora, batch.seq_off, p)
                if ea != oa or eb != ob:
                    k = next((i for i in range(min(len(ea), len(oa))) if ea[i] != oa[i] or eb[i] != ob[i]), min(len(ea), len(oa)))
                    bad = "strings differ (len %d vs %d) first at %d: eng %r/%r ora %r/%r" % (
                        len(ea), len(oa), k, ea[max(0, k - 8):k + 8], eb[max(0, k - 8):k + 8], oa[max(0, k - 8):k + 8], ob[max(0, k - 8):k + 8])
        if bad:
            n_bad += 1
            if len(msgs) < max_report:
                s0, s1, s2 = (int(batch.seq_off[2 * p]), int(batch.seq_off[2 * p + 1]), int(batch.seq_off[2 * p + 2]))
                msgs.append("%s pair %d (n=%d m=%d): %s" % (label, p, s1 - s0, s2 - s1, bad))
    if n_bad:
        msgs.append("%s %d / %d pairs differ" % (label, n_bad, batch.n_pairs))
    return msgs
