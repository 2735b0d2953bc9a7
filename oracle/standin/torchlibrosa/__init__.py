"""ORACLE / TEST INFRASTRUCTURE ONLY -- not product code.

Stand-in for the third-party package ``torchlibrosa==0.1.0`` (reference
``requirements.txt:7``), which is not installed in this image and whose source
is not under /root/reference.  Only the three classes the reference imports at
``mellow/model/htsat.py:7-8`` are provided; their arithmetic restates the
package's published algorithm.
"""
