"""Alias of pioneer_b200 under the reference's package name; see compat/README.md."""
