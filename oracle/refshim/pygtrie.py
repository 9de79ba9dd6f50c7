"""TEST INFRASTRUCTURE ONLY -- stand-in for the third-party ``pygtrie`` package (>=2.1,<3.0).

The reference imports ``pygtrie.CharTrie`` unconditionally (reference
``pyctcdecode/language_model.py:13``) and the package is not installed in this image.
Only the calls the reference makes are provided (``language_model.py:135,145,183,188,263,331``):
``CharTrie()``, ``CharTrie.fromkeys(keys)``, ``has_node(key)``, ``iterkeys(prefix, shallow=True)``.

Semantics restated from the published pygtrie behaviour: children are kept in insertion
order, ``has_node`` is truthy when the key is a stored key or a proper prefix of one, and
``iterkeys(prefix, shallow=True)`` walks the sub-trie depth first in child insertion order
and does not descend below a node that holds a value.  Never imported by the product.
"""


class _Node:
    __slots__ = ("children", "has_value")

    def __init__(self):
        self.children = {}
        self.has_value = False


class CharTrie:
    def __init__(self):
        self._root = _Node()

    @classmethod
    def fromkeys(cls, keys, value=None):
        trie = cls()
        for key in keys:
            trie[key] = value
        return trie

    def __setitem__(self, key, value):
        node = self._root
        for ch in key:
            nxt = node.children.get(ch)
            if nxt is None:
                nxt = node.children[ch] = _Node()
            node = nxt
        node.has_value = True

    def _find(self, key):
        node = self._root
        for ch in key:
            node = node.children.get(ch)
            if node is None:
                return None
        return node

    def has_node(self, key):
        node = self._find(key)
        if node is None:
            return 0
        # pygtrie returns HAS_VALUE(1) | HAS_SUBTRIE(2)
        return int(node.has_value) | (2 if node.children else 0)

    def has_key(self, key):
        node = self._find(key)
        return node is not None and node.has_value

    __contains__ = has_key

    def iterkeys(self, prefix="", shallow=False):
        node = self._find(prefix)
        if node is None:
            raise KeyError(prefix)
        stack = [(prefix, node)]
        while stack:
            key, cur = stack.pop()
            if cur.has_value:
                yield key
                if shallow:
                    continue
            for ch, child in reversed(list(cur.children.items())):
                stack.append((key + ch, child))
