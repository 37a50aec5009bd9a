"""Test-infrastructure package: CPU oracle + reference-built checkers (see dsstne_oracle.h)."""
