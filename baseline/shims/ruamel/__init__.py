# Harness-only stand-in package so the read-only reference can import `ruamel.yaml`
# in containers where ruamel.yaml is not installed.  NOT product code.
