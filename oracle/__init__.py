"""ORACLE — TEST INFRASTRUCTURE ONLY (CPU restatement of the reference hot path).  See oracle.hpp."""
