"""rails_b200 — B200-native (sm_100a) engine for ONE hot path of bailuding/rails: exact brute-force
Mixture-of-Logits top-k retrieval (rails.indexing.mol_top_k.MoLBruteForceTopK ->
rails.similarities.mol.similarity_fn.MoLSimilarity).  See DESIGN.md / INTEGRATION.md.
"""
__version__ = "0.1.0"
