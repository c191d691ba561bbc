"""TEST INFRASTRUCTURE: CPU restatement of the reference algorithm and access to the
reference CUDA build (oracle/_ref).  Never imported by the product package."""
