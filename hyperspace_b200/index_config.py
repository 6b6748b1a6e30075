"""IndexConfig -- mirror of ``python/hyperspace/indexconfig.py`` and
``src/main/scala/com/microsoft/hyperspace/index/covering/CoveringIndexConfig.scala:37-151`` (validation:
``CoveringIndexConfigTrait.scala:23-75``).  ``numBuckets`` is NOT a constructor argument: it comes from the session conf
``spark.hyperspace.index.numBuckets`` (default 200) and is frozen into the log entry at creation time."""
from typing import List, Sequence


class CoveringIndexConfig:
    def __init__(self, indexName: str, indexedColumns: Sequence[str], includedColumns: Sequence[str] = ()):
        self.indexName = indexName
        self.indexedColumns: List[str] = list(indexedColumns)
        self.includedColumns: List[str] = list(includedColumns)
        self._validate()

    def _validate(self):
        if not self.indexName or not self.indexName.strip():
            raise ValueError("Empty index name is not allowed.")
        if not self.indexedColumns:
            raise ValueError("Empty indexed columns are not allowed.")
        li = [c.lower() for c in self.indexedColumns]
        lc = [c.lower() for c in self.includedColumns]
        if len(set(li)) < len(li):
            raise ValueError("Duplicate indexed column names are not allowed.")
        if len(set(lc)) < len(lc):
            raise ValueError("Duplicate included column names are not allowed.")
        if set(li) & set(lc):
            raise ValueError("Duplicate column names in indexed/included columns are not allowed.")

    @property
    def referencedColumns(self) -> List[str]:
        return self.indexedColumns + self.includedColumns

    def __eq__(self, o):
        return (isinstance(o, CoveringIndexConfig) and self.indexName.lower() == o.indexName.lower()
                and [c.lower() for c in self.indexedColumns] == [c.lower() for c in o.indexedColumns]
                and sorted(c.lower() for c in self.includedColumns) == sorted(c.lower() for c in o.includedColumns))

    def __hash__(self):
        return hash((self.indexName.lower(), tuple(c.lower() for c in self.indexedColumns),
                     frozenset(c.lower() for c in self.includedColumns)))

    def __repr__(self):
        return (f"[indexName: {self.indexName}; indexedColumns: {','.join(self.indexedColumns)}; "
                f"includedColumns: {','.join(self.includedColumns)}]")


    # ---- builder (CoveringIndexConfig.scala:66-151): IndexConfig.builder().indexName("n").indexBy("a").include("b").create()
    class Builder:
        def __init__(self):
            self._name, self._indexed, self._included = "", [], []

        def indexName(self, indexName: str) -> "CoveringIndexConfig.Builder":
            if self._name:
                raise NotImplementedError("Index name is already set.")  # UnsupportedOperationException in the reference
            if not indexName:
                raise ValueError("Empty index name is not allowed.")
            self._name = indexName
            return self

        def indexBy(self, indexedColumn: str, *indexedColumns: str) -> "CoveringIndexConfig.Builder":
            if self._indexed:
                raise NotImplementedError("Indexed columns are already set.")
            self._indexed = [indexedColumn, *indexedColumns]
            return self

        def include(self, includedColumn: str, *includedColumns: str) -> "CoveringIndexConfig.Builder":
            if self._included:
                raise NotImplementedError("Included columns are already set.")
            self._included = [includedColumn, *includedColumns]
            return self

        def create(self) -> "CoveringIndexConfig":
            return CoveringIndexConfig(self._name, self._indexed, self._included)

    @staticmethod
    def builder() -> "CoveringIndexConfig.Builder":
        return CoveringIndexConfig.Builder()


IndexConfig = CoveringIndexConfig  # S/index/package.scala:27-33 keeps the old name as an alias
