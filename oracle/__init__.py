"""CPU oracle for the vector-mix -> flat inner-product top-k -> TREC path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as
the checker (or as the timed CPU baseline), never as the thing shipped.

What it restates (all citations are into the upstream reference tree):

* ``mix.py``     -- ``safe_mix``  onepass_dense_mix_run_custom_lang.py:342-377
                    (identical copy onepass_bilingual_mix_hub_custom_lang.py:390-424)
* ``search.py``  -- ``faiss.IndexIDMap(faiss.IndexFlatIP(d)).search`` semantics as
                    used at onepass_dense_mix_run_custom_lang.py:878 and
                    onepass_bilingual_mix_hub_custom_lang.py:950.  The arithmetic
                    lives in faiss-gpu 1.8.0 (conda-forge; pinned only in the
                    reference README.md:36,46-49), which is NOT vendored and NOT
                    installable here, so this part restates FAISS's published
                    algorithm (blocked fp32 SGEMM + k-selection, results sorted
                    descending, -1 / lowest-float padding).
* ``trec.py``    -- TREC line formatting onepass_dense_mix_run_custom_lang.py:879-888,
                    raw bilingual lines + ``collapse_run_max``
                    onepass_bilingual_mix_hub_custom_lang.py:942-962,165-181,
                    ``format_alpha`` / ``parse_alpha_list`` :304-308 / :287-301.

Pinning status
--------------
* mix / format_alpha / parse_alpha_list / collapse_run_max / TREC f-strings are
  PINNED: ``tests/golden/make_golden.py`` executes the reference's own function
  bodies (extracted from /root/reference with ``ast`` at generation time, never
  copied into this repo) on seeded inputs and the outputs are committed under
  ``tests/golden/``; ``tests/test_oracle_golden.py`` checks the oracle against them.
* the FAISS search itself is **parity unpinned**: the reference ships no tests,
  no golden vectors and no stored runs, and faiss cannot be imported here.  The
  search oracle is validated against an fp64 brute-force truth instead.
"""

from .mix import mix_normalize, mix_normalize_f64  # noqa: F401
from .search import (  # noqa: F401
    flat_ip_search,
    flat_ip_search_f64,
    merge_topk,
    compare_topk,
)
from .trec import (  # noqa: F401
    format_alpha,
    parse_alpha_list,
    mono_trec_lines,
    bilingual_raw_lines,
    collapse_run_max,
    collapse_run_max_text,
)
