"""Stub of Biopython for importing the reference in the authoring container (test infrastructure only)."""
