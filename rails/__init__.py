"""`rails` import alias of the B200-native package (the replacement classes must be importable under the
reference's own module paths, e.g. `rails.indexing.mol_top_k.MoLBruteForceTopK`,
`rails.similarities.mol.similarity_fn.MoLSimilarity`).

Every `rails.X` below IS the module `rails_b200.X` (same module object, registered under both names), so isinstance
checks and state held on the modules agree whichever path a caller imports.  Put this repo on sys.path INSTEAD of the
reference's `rails/` directory; the reference's out-of-path subsystems (training, datasets, encoders) are not here.
"""
import importlib
import sys

_MODULES = (
    "indexing", "indexing.candidate_index", "indexing.mol_top_k", "indexing.mips_top_k",
    "similarities", "similarities.module", "similarities.layers", "similarities.dot_product_similarity_fn",
    "similarities.mol", "similarities.mol.embeddings_fn", "similarities.mol.item_embeddings_fns",
    "similarities.mol.query_embeddings_fns", "similarities.mol.similarity_fn",
)
for _name in _MODULES:
    _mod = importlib.import_module("rails_b200." + _name)
    sys.modules["rails." + _name] = _mod
    if "." not in _name:
        globals()[_name] = _mod
del _name, _mod
