"""`said` import-name shim over said_b200 (see compat/README.md)."""
