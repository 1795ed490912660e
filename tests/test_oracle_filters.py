"""Column filters: the oracle's restatement of ApplyFiltersWithBitmap / ApplyFilters
(src/server/search_pipeline.cpp:1021-1237) against the reference's own DocumentStore + FilterIndex + pipeline code
(oracle/_ref), and against the known answers of tests/server/search_pipeline_test.cpp:71-395."""
import random

import numpy as np

INT_TYPES = {2: (-128, 127), 3: (0, 255), 4: (-32768, 32767), 5: (0, 65535), 6: (-2 ** 31, 2 ** 31 - 1),
             7: (0, 2 ** 32 - 1), 8: (-2 ** 63, 2 ** 63 - 1), 9: (0, 2 ** 64 - 1), 10: (-86400, 86400)}
LITERALS = ["1", "0", "true", "false", "5", "-5", "05", "5.0", "2.5", "-0", "0.0", "abc", "", "b", "ab", "300", "70000",
            "4294967296", "18446744073709551615", "-9223372036854775808", "1e3", "nan", " 5", "5 ", "TRUE"]


def random_columns(rnd, n_docs):
    cols = []
    for typ in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12):
        vals = []
        for _ in range(n_docs):
            r = rnd.random()
            if r < 0.15:
                vals.append(None)
            elif typ == 1:
                vals.append(rnd.random() < 0.5)
            elif typ in INT_TYPES:
                lo, hi = INT_TYPES[typ]
                vals.append(rnd.choice([0, 1, 5, -5, 300, 70000, lo, hi, rnd.randint(lo, hi)]) if lo < 0 else
                            rnd.choice([0, 1, 5, 300, 70000, hi, rnd.randint(lo, hi)]))
                vals[-1] = max(lo, min(hi, vals[-1]))
            elif typ == 11:
                vals.append(rnd.choice([b"abc", b"ab", b"b", b"", b"5", b"true", b"zz", "東京".encode()]))
            else:
                vals.append(rnd.choice([0.0, -0.0, 1.0, 2.5, 5.0, 5.0 + 1e-12, 1000.0, -5.0, 1e300]))
        cols.append((typ, vals))
    return cols


def test_filter_restatement_matches_reference_sources(oracle, reflib):
    rnd = random.Random(123)
    n_docs = 60
    for trial in range(40):
        cols = random_columns(rnd, n_docs)
        first = rnd.choice([1, 100])
        results = sorted(rnd.sample(range(first, first + n_docs), rnd.randint(0, n_docs)))
        for _ in range(25):
            n_f = rnd.randint(1, 3)
            eq_only = rnd.random() < 0.5
            filters = [(rnd.randrange(len(cols) + (1 if rnd.random() < 0.05 else 0)),
                        rnd.choice([0, 1]) if eq_only else rnd.randrange(6), rnd.choice(LITERALS)) for _ in range(n_f)]
            a = oracle.apply_filters(n_docs, first, cols, filters, results)
            b = reflib.apply_filters(n_docs, first, cols, filters, results)
            assert np.array_equal(a, b), (filters, a, b)


def test_filter_known_answers(oracle):
    """tests/server/search_pipeline_test.cpp: ApplyFilters with int / string columns and NULLs."""
    # docs 1..4: status int32 = 1, 2, NULL, 1 ; category string = "a", "b", "a", NULL
    cols = [(6, [1, 2, None, 1]), (11, [b"a", b"b", b"a", None])]
    allr = [1, 2, 3, 4]
    assert oracle.apply_filters(4, 1, cols, [(0, 0, "1")], allr).tolist() == [1, 4]          # status = 1
    assert oracle.apply_filters(4, 1, cols, [(0, 1, "1")], allr).tolist() == [2, 3]          # != keeps NULL
    assert oracle.apply_filters(4, 1, cols, [(0, 2, "1")], allr).tolist() == [2]             # > 1 (NULL never matches)
    assert oracle.apply_filters(4, 1, cols, [(1, 0, "a"), (0, 0, "1")], allr).tolist() == [1]
    assert oracle.apply_filters(4, 1, cols, [(1, 5, "a")], allr).tolist() == [1, 3]          # <= "a"
    assert oracle.apply_filters(4, 1, cols, [(0, 0, "abc")], allr).tolist() == []            # not a number
