"""fasthigashi_b200 - B200-native implementation of Fast-Higashi's decomposition hot path
(partial RWR imputation feeding the integrative PARAFAC2 ALS loop). See DESIGN.md."""
__version__ = "0.1.0"
