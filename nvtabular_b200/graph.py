"""Operator-graph sugar of the reference API: ColumnSelector, Node (``>>``,
``+``, ``-``, ``[]``), a minimal Schema.  In the reference these are aliases
of un-vendored merlin.dag / merlin.schema classes (nvtabular/graph.py:21,
nvtabular/workflow/node.py:16-18, nvtabular/ops/operator.py:16-27); only the
surface the hot-path operators and Workflow.fit/transform use is rebuilt here
(SURVEY.md §8b), exercised by reference tests/unit/workflow/test_workflow_node.py.
"""
from enum import Enum
from typing import Dict, List, Optional, Union


class Tags(Enum):
    CATEGORICAL = "categorical"
    CONTINUOUS = "continuous"
    LIST = "list"


class ColumnSelector:
    """names + nested groups: ``ColumnSelector(["a", ["b", "c"]])`` selects a, b, c
    and remembers that (b, c) is a multi-column group."""

    def __init__(self, names=None, subgroups=None):
        self._names: List[str] = []
        self.subgroups: List["ColumnSelector"] = list(subgroups or [])
        if names is None:
            names = []
        if isinstance(names, str):
            names = [names]
        for name in names:
            if isinstance(name, str):
                self._names.append(name)
            elif isinstance(name, ColumnSelector):
                self.subgroups.append(name)
            elif isinstance(name, (list, tuple)):
                self.subgroups.append(ColumnSelector(list(name)))
            elif isinstance(name, Node):
                raise TypeError("Nodes can not be passed to the constructor of ColumnSelector")
            else:
                raise TypeError(f"unsupported column name {name!r}")
        for sg in self.subgroups:
            if sg.subgroups:
                raise ValueError("nested subgroups are not allowed")

    @property
    def names(self) -> List[str]:
        out = list(self._names)
        for sg in self.subgroups:
            out += sg.names
        return out

    @property
    def grouped_names(self) -> List[Union[str, tuple]]:
        out: List[Union[str, tuple]] = list(self._names)
        for sg in self.subgroups:
            out.append(tuple(sg.names))
        return out

    def __add__(self, other):
        if other is None:
            return self
        if isinstance(other, Node):
            return other.__radd__(self)
        if isinstance(other, (str, list, tuple)):
            other = ColumnSelector(other)
        return ColumnSelector(self._names + other._names, self.subgroups + other.subgroups)

    def __radd__(self, other):
        return ColumnSelector(other) + self

    def __rshift__(self, operator):
        return Node(self) >> operator

    def __eq__(self, other):
        if not isinstance(other, ColumnSelector):
            return False
        return self._names == other._names and self.subgroups == other.subgroups

    def __bool__(self):
        return bool(self._names or self.subgroups)

    def __repr__(self):
        return f"ColumnSelector({self.grouped_names})"


class ColumnSchema:
    def __init__(self, name, dtype=None, tags=None, properties=None, is_list=False, is_ragged=False):
        self.name = name
        self.dtype = dtype
        self.tags = list(tags or [])
        self.properties = dict(properties or {})
        self.is_list = is_list
        self.is_ragged = is_ragged

    def with_name(self, name):
        return ColumnSchema(name, self.dtype, self.tags, self.properties, self.is_list, self.is_ragged)

    def with_dtype(self, dtype, is_list=None, is_ragged=None):
        return ColumnSchema(self.name, dtype, self.tags, self.properties,
                            self.is_list if is_list is None else is_list,
                            self.is_ragged if is_ragged is None else is_ragged)

    def with_tags(self, tags):
        merged = list(dict.fromkeys(self.tags + list(tags)))
        return ColumnSchema(self.name, self.dtype, merged, self.properties, self.is_list, self.is_ragged)

    def with_properties(self, props):
        return ColumnSchema(self.name, self.dtype, self.tags, {**self.properties, **props},
                            self.is_list, self.is_ragged)

    def __repr__(self):
        return f"ColumnSchema({self.name!r}, dtype={self.dtype}, tags={self.tags})"


class Schema:
    def __init__(self, column_schemas=None):
        self.column_schemas: Dict[str, ColumnSchema] = {}
        for cs in (column_schemas or []):
            if isinstance(cs, str):
                cs = ColumnSchema(cs)
            self.column_schemas[cs.name] = cs

    @property
    def column_names(self):
        return list(self.column_schemas)

    def __getitem__(self, name):
        if isinstance(name, (list, tuple)):
            return Schema([self.column_schemas[n] for n in name])
        return self.column_schemas[name]

    def __contains__(self, name):
        return name in self.column_schemas

    def __iter__(self):
        return iter(self.column_schemas.values())

    def __len__(self):
        return len(self.column_schemas)

    def __bool__(self):
        return True

    def __add__(self, other):
        return Schema(list(self.column_schemas.values()) + list(other.column_schemas.values()))

    def select_by_tag(self, tag):
        return Schema([c for c in self if tag in c.tags])

    def select_by_name(self, names):
        return Schema([self.column_schemas[n] for n in names if n in self.column_schemas])

    def without(self, names):
        return Schema([c for c in self if c.name not in set(names)])


class Node:
    """One vertex of the operator DAG."""

    def __init__(self, selector: Optional[ColumnSelector] = None):
        self.parents: List["Node"] = []
        self.children: List["Node"] = []
        self.dependencies: List["Node"] = []
        self.op = None
        self.selector = selector
        self.kind = "input" if selector is not None else "op"   # input | op | concat | subtract | subset
        self.input_schema: Optional[Schema] = None
        self.output_schema: Optional[Schema] = None

    # ---------------------------------------------------------------- building
    @classmethod
    def construct_from(cls, obj) -> "Node":
        if isinstance(obj, Node):
            return obj
        if isinstance(obj, ColumnSelector):
            return Node(obj)
        if isinstance(obj, str):
            return Node(ColumnSelector([obj]))
        if isinstance(obj, (list, tuple)):
            if all(isinstance(x, str) or isinstance(x, (list, tuple)) and all(isinstance(y, str) for y in x)
                   for x in obj):
                return Node(ColumnSelector(list(obj)))
            node = Node()
            node.kind = "concat"
            for x in obj:
                node.add_parent(cls.construct_from(x))
            return node
        raise TypeError(f"cannot build a graph node from {type(obj)}")

    def add_parent(self, parent: "Node"):
        self.parents.append(parent)
        parent.children.append(self)

    def add_dependency(self, dep):
        dep = Node.construct_from(dep)
        self.dependencies.append(dep)
        dep.children.append(self)

    def __rshift__(self, operator):
        from .ops.base import Operator
        if isinstance(operator, type) and issubclass(operator, Operator):
            operator = operator()
        if not isinstance(operator, Operator):
            raise ValueError(f"Expected operator or callable, got {type(operator)}")
        child = Node()
        child.op = operator
        child.add_parent(self)
        deps = operator.dependencies
        if deps:
            for d in (deps if isinstance(deps, list) else [deps]):
                child.add_dependency(d)
        return child

    def __rrshift__(self, other):
        return ColumnSelector(other) >> self

    def __add__(self, other):
        node = Node()
        node.kind = "concat"
        node.add_parent(self)
        others = other if isinstance(other, list) and any(isinstance(o, Node) for o in other) else [other]
        for o in others:
            node.add_parent(Node.construct_from(o))
        return node

    def __radd__(self, other):
        node = Node()
        node.kind = "concat"
        node.add_parent(Node.construct_from(other))
        node.add_parent(self)
        return node

    def __sub__(self, other):
        node = Node()
        node.kind = "subtract"
        node.add_parent(self)
        node.selector = other if isinstance(other, ColumnSelector) else ColumnSelector(
            other.output_columns.names if isinstance(other, Node) else other)
        node._subtract_node = other if isinstance(other, Node) else None
        return node

    def __getitem__(self, columns):
        node = Node()
        node.kind = "subset"
        node.add_parent(self)
        node.selector = columns if isinstance(columns, ColumnSelector) else ColumnSelector(columns)
        return node

    # --------------------------------------------------------------- traversal
    @property
    def upstream(self) -> List["Node"]:
        return self.parents + self.dependencies

    def topo_order(self) -> List["Node"]:
        seen, order = set(), []

        def visit(n):
            if id(n) in seen:
                return
            seen.add(id(n))
            for p in n.upstream:
                visit(p)
            order.append(n)
        visit(self)
        return order

    # ----------------------------------------------------------------- columns
    @property
    def input_columns(self) -> ColumnSelector:
        """Selector handed to op.fit/op.transform: the parents' output columns,
        keeping the grouping of selector parents (multi-column groups)."""
        names, subgroups = [], []
        for p in self.parents:
            sel = p.output_columns
            for n in sel._names:
                if n not in names:
                    names.append(n)
            subgroups += sel.subgroups
        return ColumnSelector(names, subgroups)

    @property
    def output_columns(self) -> ColumnSelector:
        if self.kind == "input":
            return self.selector
        if self.kind == "concat":
            return self.input_columns
        if self.kind == "subtract":
            drop = set(self.selector.names)
            inp = self.input_columns
            return ColumnSelector([n for n in inp._names if n not in drop],
                                  [sg for sg in inp.subgroups if not set(sg.names) & drop])
        if self.kind == "subset":
            return self.selector
        return ColumnSelector(list(self.op.column_mapping(self.input_columns).keys()))

    @property
    def dependency_columns(self) -> ColumnSelector:
        names = []
        for d in self.dependencies:
            names += [n for n in d.output_columns.names if n not in names]
        return ColumnSelector(names)

    def root_columns(self) -> List[str]:
        """Names of the raw dataset columns this (sub)graph reads."""
        out = []
        for n in self.topo_order():
            if n.kind == "input":
                out += [c for c in n.selector.names if c not in out]
        return out

    def __repr__(self):
        if self.kind == "input":
            return f"<Node input {self.selector.grouped_names}>"
        if self.kind == "op":
            return f"<Node {type(self.op).__name__}>"
        return f"<Node {self.kind}>"
