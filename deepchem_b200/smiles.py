"""SMILES -> ``ConvMol`` without RDKit: a reader for the SMILES the reference's GraphConv
datasets hold (organic subset, bracket atoms with H counts / charges / isotopes, aromatic lower case, ring closures
incl. ``%nn``, ``- = # : / \\ .``) and the 75-dim atom feature vector of ``deepchem/feat/graph_features.py:282-391``
(``atom_features``) as ``ConvMolFeaturizer._featurize`` (``:845-914``) assembles it:

    [0:44)  element one-hot (unknown -> 43)        [44:55) heavy-atom degree 0..10
    [55:62) implicit valence 0..6 (unk -> 6)       62 formal charge        63 radical electrons
    [64:69) hybridization SP, SP2, SP3, SP3D, SP3D2 (anything else -> SP3D2, ``one_of_k_encoding_unk``)
    69      aromatic                               [70:75) total hydrogens 0..4 (unk -> 4)

RDKit supplies those atom properties in the reference; here they are derived with the published rules RDKit's
sanitisation follows (restated below next to each step): valence model with charge-shifted default valences,
Kekulé assignment of lower-case rings by perfect matching, Hueckel perception of rings written in Kekulé form,
conjugation and hybridization from bonds + lone pairs.  PARITY of this module is pinned only by the reference's
known answer for ``['CCC', 'C']`` (``tests/golden/kat_ccc_c.npz``) and, for hydrogen counts and ring
counts over all 1 128 Delaney molecules, by the descriptor columns the reference's own
``datasets/delaney-processed.csv`` carries (molecular weight, H-bond donors, rings, minimum degree) — RDKit itself
is absent from this image.  Known differences: atoms are numbered in SMILES order (the reference renumbers them by
RDKit's canonical ranking, ``feat/base_classes.py:305-308``), and the order of ring-closure bonds in the adjacency
lists may differ; GraphConv models are invariant to both up to floating-point summation order.  ``use_chirality``
and ``atom_properties`` need RDKit's CIP perception / property tables and raise ``NotImplementedError``.
"""
import numpy as np

from .mol_graphs import ConvMol

# element -> (atomic number, outer-shell electrons, RDKit valence list (-1 = unconstrained), atomic weight)
_ELEMENTS = {
    '*': (0, 0, (-1,), 0.0), 'H': (1, 1, (1,), 1.008), 'He': (2, 2, (0,), 4.003), 'Li': (3, 1, (1, -1), 6.941),
    'Be': (4, 2, (2,), 9.012), 'B': (5, 3, (3,), 10.812), 'C': (6, 4, (4,), 12.011), 'N': (7, 5, (3,), 14.007),
    'O': (8, 6, (2,), 15.999), 'F': (9, 7, (1,), 18.998), 'Ne': (10, 8, (0,), 20.18), 'Na': (11, 1, (1, -1), 22.99),
    'Mg': (12, 2, (2, -1), 24.305), 'Al': (13, 3, (3, 6), 26.982), 'Si': (14, 4, (4, 6), 28.086),
    'P': (15, 5, (3, 5, 7), 30.974), 'S': (16, 6, (2, 4, 6), 32.067), 'Cl': (17, 7, (1,), 35.453),
    'Ar': (18, 8, (0,), 39.948), 'K': (19, 1, (1, -1), 39.098), 'Ca': (20, 2, (2, -1), 40.078),
    'Ti': (22, 4, (-1,), 47.867), 'V': (23, 5, (-1,), 50.942), 'Cr': (24, 6, (-1,), 51.996),
    'Mn': (25, 7, (-1,), 54.938), 'Fe': (26, 8, (-1,), 55.845), 'Co': (27, 9, (-1,), 58.933),
    'Ni': (28, 10, (-1,), 58.693), 'Cu': (29, 11, (-1,), 63.546), 'Zn': (30, 2, (-1,), 65.39),
    'Ge': (32, 4, (4,), 72.61), 'As': (33, 5, (3, 5, 7), 74.922), 'Se': (34, 6, (2, 4, 6), 78.96),
    'Br': (35, 7, (1,), 79.904), 'Zr': (40, 4, (-1,), 91.224), 'Pd': (46, 10, (-1,), 106.42),
    'Ag': (47, 11, (-1,), 107.868), 'Cd': (48, 2, (-1,), 112.412), 'In': (49, 3, (3,), 114.818),
    'Sn': (50, 4, (2, 4), 118.711), 'Sb': (51, 5, (3, 5, 7), 121.76), 'Te': (52, 6, (2, 4, 6), 127.6),
    'I': (53, 7, (1, 3, 5), 126.904), 'Yb': (70, 3, (-1,), 173.04), 'Pt': (78, 10, (-1,), 195.078),
    'Au': (79, 11, (-1,), 196.967), 'Hg': (80, 2, (-1,), 200.59), 'Tl': (81, 3, (-1,), 204.383),
    'Pb': (82, 4, (2, 4), 207.2),
    # heavy elements of the reference's Tox21 CSV (all land in the 'Unknown' element slot)
    'Sr': (38, 2, (2, -1), 87.62), 'Ba': (56, 2, (2, -1), 137.327), 'Nd': (60, 4, (-1,), 144.24),
    'Bi': (83, 5, (3, 5, 7), 208.98),
}
# every two-letter element symbol, so that '[Ba+2]' is not read as boron followed by garbage; symbols outside
# _ELEMENTS are accepted as 'Unknown' atoms with unconstrained valence
_TWO_LETTER = frozenset(
    'He Li Be Ne Na Mg Al Si Cl Ar Ca Sc Ti Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr Rb Sr Zr Nb Mo Tc Ru Rh Pd Ag Cd In '
    'Sn Sb Te Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac '
    'Th Pa Np Pu Am Cm Bk Cf Es Fm Md No Lr Rf Db Sg Bh Hs Mt Ds Rg Cn Nh Fl Mc Lv Ts Og'.split())
# graph_features.py:322-367
_SYMBOLS = ['C', 'N', 'O', 'S', 'F', 'Si', 'P', 'Cl', 'Br', 'Mg', 'Na', 'Ca', 'Fe', 'As', 'Al', 'I', 'B', 'V', 'K', 'Tl',
            'Yb', 'Sb', 'Sn', 'Ag', 'Pd', 'Co', 'Se', 'Ti', 'Zn', 'H', 'Li', 'Ge', 'Cu', 'Au', 'Ni', 'Cd', 'In', 'Mn', 'Zr',
            'Cr', 'Pt', 'Hg', 'Pb', 'Unknown']
_SYMBOL_INDEX = {s: i for i, s in enumerate(_SYMBOLS)}
_ORGANIC = ('Cl', 'Br', 'B', 'C', 'N', 'O', 'P', 'S', 'F', 'I')
_AROMATIC = {'b': 'B', 'c': 'C', 'n': 'N', 'o': 'O', 'p': 'P', 's': 'S', 'se': 'Se', 'as': 'As', 'te': 'Te'}
_EARLY = {1, 3, 4, 5, 11, 12, 13, 19, 20}       # charge raises (not lowers) the usable valence: B-, Al-
_BOND_ORDER = {'-': 1, '/': 1, '\\': 1, '=': 2, '#': 3, '$': 4, ':': 1}
S, SP, SP2, SP3, SP3D, SP3D2, OTHER = range(7)
N_ATOM_FEATURES = 75


class SmilesError(ValueError):
    """The string is not a SMILES this reader accepts (RDKit's ``MolFromSmiles`` returning None)."""


class Atom(object):
    __slots__ = ("symbol", "aromatic", "bracket", "hcount", "charge", "isotope", "nbrs", "bonds", "implicit_h",
                 "radicals", "hybridization", "in_ring")

    def __init__(self, symbol, aromatic=False, bracket=False, hcount=0, charge=0, isotope=0):
        self.symbol, self.aromatic, self.bracket = symbol, aromatic, bracket
        self.hcount, self.charge, self.isotope = hcount, charge, isotope        # hcount: hydrogens written in brackets
        self.nbrs, self.bonds = [], []
        self.implicit_h, self.radicals, self.hybridization, self.in_ring = 0, 0, OTHER, False

    z = property(lambda self: _ELEMENTS.get(self.symbol, _ELEMENTS['*'])[0])
    n_outer = property(lambda self: _ELEMENTS.get(self.symbol, _ELEMENTS['*'])[1])
    valences = property(lambda self: _ELEMENTS.get(self.symbol, _ELEMENTS['*'])[2])
    mass = property(lambda self: _ELEMENTS.get(self.symbol, _ELEMENTS['*'])[3])
    total_h = property(lambda self: self.hcount + self.implicit_h)
    degree = property(lambda self: len(self.nbrs))


class Bond(object):
    __slots__ = ("a", "b", "order", "aromatic", "in_ring", "conjugated", "explicit")

    def __init__(self, a, b, order, aromatic, explicit):
        self.a, self.b, self.order, self.aromatic, self.explicit = a, b, order, aromatic, explicit
        self.in_ring = self.conjugated = False

    def other(self, i):
        return self.b if i == self.a else self.a


class Mol(object):
    """Atoms in SMILES order; bonds: chain bonds in the order written, then ring-closure bonds in closing order."""

    def __init__(self, atoms, bonds):
        self.atoms, self.bonds = atoms, bonds

    def adjacency_list(self):
        """``ConvMolFeaturizer._featurize``: for every bond, end atom appended to the begin atom's list and vice versa
        (graph_features.py:896-904)."""
        adj = [[] for _ in self.atoms]
        for b in self.bonds:
            adj[b.a].append(b.b)
            adj[b.b].append(b.a)
        return adj

    def molecular_weight(self):
        return sum(a.mass + _ELEMENTS['H'][3] * a.total_h for a in self.atoms)

    def n_rings(self):
        """Cyclomatic number (bonds - atoms + fragments) = the size of RDKit's SSSR."""
        parent = list(range(len(self.atoms)))

        def find(i):
            while parent[i] != i:
                parent[i] = parent[parent[i]]
                i = parent[i]
            return i
        for b in self.bonds:
            parent[find(b.a)] = find(b.b)
        return len(self.bonds) - len(self.atoms) + len({find(i) for i in range(len(self.atoms))})


# ------------------------------------------------------------------------------------------------ parsing
def _parse_bracket(s, i):
    """``[`` isotope? symbol chirality? H-count? charge? class? ``]`` starting after the ``[``; returns (Atom, next)."""
    n = len(s)
    j = i
    while j < n and s[j].isdigit():
        j += 1
    isotope = int(s[i:j]) if j > i else 0
    if j >= n:
        raise SmilesError("unterminated bracket atom")
    aromatic = False
    if s[j:j + 2] in _AROMATIC and s[j:j + 2] in ('se', 'as', 'te'):
        symbol, aromatic, j = _AROMATIC[s[j:j + 2]], True, j + 2
    elif s[j] in _AROMATIC:
        symbol, aromatic, j = _AROMATIC[s[j]], True, j + 1
    elif s[j] == '*':
        symbol, j = '*', j + 1
    elif s[j].isupper():
        if j + 1 < n and s[j + 1].islower() and (s[j:j + 2] in _ELEMENTS or s[j:j + 2] in _TWO_LETTER):
            symbol, j = s[j:j + 2], j + 2                                   # (outside _ELEMENTS: an 'Unknown' atom)
        elif j + 1 < n and s[j + 1].islower() and s[j + 1] not in 'h' and s[j] not in _ELEMENTS:
            symbol, j = s[j:j + 2], j + 2                                   # not a symbol at all: 'Unknown' as well
        else:
            symbol, j = s[j], j + 1
    else:
        raise SmilesError("bad bracket atom at %d" % j)
    while j < n and s[j] == '@':                                            # chirality marks (not used by the 75 features)
        j += 1
    if s[j:j + 2] in ('TH', 'AL', 'SP', 'TB', 'OH'):
        j += 2
        while j < n and s[j].isdigit():
            j += 1
    hcount = 0
    if j < n and s[j] == 'H':
        j += 1
        k = j
        while j < n and s[j].isdigit():
            j += 1
        hcount = int(s[k:j]) if j > k else 1
    charge = 0
    if j < n and s[j] in '+-':
        sign = 1 if s[j] == '+' else -1
        k = j
        while j < n and s[j] == s[k]:
            j += 1
        charge = sign * (j - k)
        k = j
        while j < n and s[j].isdigit():
            j += 1
        if j > k:
            charge = sign * int(s[k:j])
    if j < n and s[j] == ':':
        j += 1
        while j < n and s[j].isdigit():
            j += 1
    if j >= n or s[j] != ']':
        raise SmilesError("bad bracket atom near %d" % j)
    return Atom(symbol, aromatic, True, hcount, charge, isotope), j + 1


def parse_smiles(smiles):
    """Atoms and bonds as written (no chemistry yet)."""
    s = smiles.strip()
    if not s:
        raise SmilesError("empty SMILES")
    atoms, chain_bonds, ring_bonds = [], [], []
    stack, prev, pending, rings = [], None, None, {}
    i, n = 0, len(s)

    def add_atom(atom):
        nonlocal prev, pending
        atoms.append(atom)
        cur = len(atoms) - 1
        if prev is not None and pending != '.':
            chain_bonds.append((prev, cur, pending))
        prev, pending = cur, None

    while i < n:
        c = s[i]
        if c == '[':
            atom, i = _parse_bracket(s, i + 1)
            add_atom(atom)
        elif c in '-=#$:/\\':
            if pending is not None:
                raise SmilesError("two bond symbols in a row at %d" % i)
            pending, i = c, i + 1
        elif c == '.':
            if prev is None or pending is not None:
                raise SmilesError("misplaced '.'")
            pending, i = '.', i + 1
        elif c == '(':
            if prev is None:
                raise SmilesError("branch before any atom")
            stack.append(prev)
            i += 1
        elif c == ')':
            if not stack or pending is not None:
                raise SmilesError("unbalanced ')' at %d" % i)
            prev, i = stack.pop(), i + 1
        elif c.isdigit() or c == '%':
            if c == '%':
                if not s[i + 1:i + 3].isdigit() or len(s[i + 1:i + 3]) != 2:
                    raise SmilesError("bad %%nn ring closure at %d" % i)
                label, i = int(s[i + 1:i + 3]), i + 3
            else:
                label, i = int(c), i + 1
            if prev is None or pending == '.':
                raise SmilesError("ring closure without an atom")
            if label in rings:
                other, sym = rings.pop(label)
                if other == prev:
                    raise SmilesError("ring closure to the same atom")
                if sym is not None and pending is not None and _BOND_ORDER[sym] != _BOND_ORDER[pending]:
                    raise SmilesError("conflicting ring closure bonds")
                ring_bonds.append((other, prev, pending if pending is not None else sym))
            else:
                rings[label] = (prev, pending)
            pending = None
        elif s[i:i + 2] in ('Cl', 'Br'):
            add_atom(Atom(s[i:i + 2]))
            i += 2
        elif c in _ORGANIC or c == '*':
            add_atom(Atom(c))
            i += 1
        elif c in 'bcnops':
            add_atom(Atom(_AROMATIC[c], aromatic=True))
            i += 1
        else:
            raise SmilesError("unexpected character %r at %d" % (c, i))
    if stack or rings or pending not in (None,):
        raise SmilesError("unclosed branch, ring or dangling bond")
    seen = set()
    bonds = []
    for a, b, sym in chain_bonds + ring_bonds:
        key = (min(a, b), max(a, b))
        if key in seen:
            raise SmilesError("duplicate bond %d-%d" % key)
        seen.add(key)
        arom = sym == ':' or (sym is None and atoms[a].aromatic and atoms[b].aromatic)
        bonds.append(Bond(a, b, _BOND_ORDER.get(sym, 1), arom, sym is not None))
    return Mol(atoms, bonds)


# ------------------------------------------------------------------------------------------------ chemistry
def _connect(mol):
    for a in mol.atoms:
        a.nbrs, a.bonds = [], []
    for k, b in enumerate(mol.bonds):
        mol.atoms[b.a].nbrs.append(b.b)
        mol.atoms[b.a].bonds.append(k)
        mol.atoms[b.b].nbrs.append(b.a)
        mol.atoms[b.b].bonds.append(k)


def _fold_hydrogens(mol):
    """RDKit's default ``removeHs``: ``[H]`` atoms with one heavy neighbour, no charge and no isotope become a
    hydrogen count on that neighbour (explicit when it is a bracket atom, implicit otherwise)."""
    _connect(mol)
    drop = set()
    for i, a in enumerate(mol.atoms):
        if a.symbol == 'H' and a.degree == 1 and not a.charge and not a.isotope and not a.hcount:
            nb = mol.atoms[a.nbrs[0]]
            if nb.symbol != 'H' and mol.bonds[a.bonds[0]].order == 1:
                drop.add(i)
                if nb.bracket:
                    nb.hcount += 1
    if not drop:
        return
    remap, atoms = {}, []
    for i, a in enumerate(mol.atoms):
        if i not in drop:
            remap[i] = len(atoms)
            atoms.append(a)
    bonds = []
    for b in mol.bonds:
        if b.a in drop or b.b in drop:
            continue
        b.a, b.b = remap[b.a], remap[b.b]
        bonds.append(b)
    mol.atoms, mol.bonds = atoms, bonds
    _connect(mol)


def _cleanup(mol):
    """RDKit ``MolOps::cleanUp`` for the one pattern GraphConv datasets are full of: a neutral five-valent nitrogen
    with a double bond to a terminal neutral oxygen (nitro ``N(=O)=O``, N-oxides ``n=O``) becomes ``[N+]-[O-]``."""
    for i, a in enumerate(mol.atoms):
        if a.symbol != 'N' or a.charge:
            continue
        valence = sum(mol.bonds[k].order for k in a.bonds) + a.hcount + (1 if a.aromatic else 0)
        if valence != 5:
            continue
        for k in a.bonds:
            b = mol.bonds[k]
            o = mol.atoms[b.other(i)]
            if o.symbol == 'O' and not o.charge and b.order == 2 and o.degree == 1:
                b.order, a.charge, o.charge = 1, 1, -1
                break


def _mark_rings(mol):
    """A bond is a ring bond iff its ends stay connected without it, i.e. iff it is not a bridge: one depth-first walk
    with low-links (Tarjan) over every fragment instead of one search per bond."""
    n = len(mol.atoms)
    order, low = [-1] * n, [0] * n
    for b in mol.bonds:
        b.in_ring = True
    clock = 0
    for root in range(n):
        if order[root] >= 0:
            continue
        order[root] = low[root] = clock
        clock += 1
        stack = [(root, -1, 0)]                       # (atom, bond it was entered by, next position in its bond list)
        while stack:
            v, via, pos = stack.pop()
            bonds = mol.atoms[v].bonds
            if pos < len(bonds):
                stack.append((v, via, pos + 1))
                k = bonds[pos]
                if k == via:
                    continue
                w = mol.bonds[k].other(v)
                if order[w] < 0:
                    order[w] = low[w] = clock
                    clock += 1
                    stack.append((w, k, 0))
                elif order[w] < low[v]:
                    low[v] = order[w]
            elif via >= 0:
                parent = mol.bonds[via].other(v)
                if low[v] < low[parent]:
                    low[parent] = low[v]
                if low[v] > order[parent]:
                    mol.bonds[via].in_ring = False    # nothing below v reaches parent or above: a bridge
    for a in mol.atoms:
        a.in_ring = any(mol.bonds[k].in_ring for k in a.bonds)


def _usable_valences(atom):
    """RDKit's valence model: the default valences shifted by the formal charge (the sign is flipped for the early
    groups, so B- takes four bonds like C, N+ four, O- one)."""
    chg = -atom.charge if atom.z in _EARLY else atom.charge
    return [v + chg for v in atom.valences if v >= 0], (-1 in atom.valences)


def _kekulize(mol):
    """Assign alternating double bonds over the bonds read as aromatic: an aromatic atom takes one double bond iff
    its usable valence leaves room for it beyond its sigma bonds and written hydrogens (c, pyridine n, n+; not
    pyrrole n / [nH], o, s, nor c with an exocyclic double bond).  Perfect matching by backtracking, most
    constrained atom first.  Raises SmilesError where RDKit reports "Can't kekulize"."""
    for b in mol.bonds:
        if b.aromatic and not b.explicit and not b.in_ring:
            b.aromatic = False                       # c1ccccc1c1ccccc1: the bond between the rings is single
    need = {}
    for i, a in enumerate(mol.atoms):
        if not a.aromatic:
            continue
        arom = [k for k in a.bonds if mol.bonds[k].aromatic]
        if not arom:
            raise SmilesError("aromatic atom %d outside a ring" % i)
        sigma = sum(mol.bonds[k].order for k in a.bonds) + a.hcount
        vals, free = _usable_valences(a)
        room = [v for v in vals if v >= sigma]
        if free and not room:
            continue
        if not room:
            raise SmilesError("valence of aromatic atom %d exceeded" % i)
        if room[0] - sigma >= 1:
            need[i] = arom
    match = {}

    def options(i):
        return [k for k in need[i] if mol.bonds[k].other(i) in need and mol.bonds[k].other(i) not in match]

    def solve():
        best, best_opts = None, None
        for i in need:
            if i in match:
                continue
            opts = options(i)
            if not opts:
                return False
            if best is None or len(opts) < len(best_opts):
                best, best_opts = i, opts
                if len(opts) == 1:
                    break
        if best is None:
            return True
        for k in best_opts:
            j = mol.bonds[k].other(best)
            match[best], match[j] = k, k
            if solve():
                return True
            del match[best], match[j]
        return False

    import sys
    limit = sys.getrecursionlimit()
    if len(need) + 100 > limit:
        sys.setrecursionlimit(len(need) + 200)
    try:
        ok = solve()
    finally:
        sys.setrecursionlimit(limit)
    if not ok:
        raise SmilesError("can't kekulize")
    for k in set(match.values()):
        mol.bonds[k].order = 2


def _assign_hydrogens(mol):
    """Implicit hydrogens of the atoms written without brackets: up to the smallest usable valence that covers the
    bond orders; radical electrons of bracket atoms (RDKit ``assignRadicals``: what is missing to the octet / to the
    lowest usable valence, never more than the unshared outer electrons)."""
    for a in mol.atoms:
        bo = sum(mol.bonds[k].order for k in a.bonds)
        vals, free = _usable_valences(a)
        if not a.bracket:
            room = [v for v in vals if v >= bo]
            if not room and not free and a.symbol != '*':
                raise SmilesError("valence of %s exceeded" % a.symbol)
            a.implicit_h = room[0] - bo if room else 0
            a.radicals = 0
        else:
            a.implicit_h = 0
            total = bo + a.hcount
            if a.z == 0:
                a.radicals = 0
            elif vals and not free:
                base = 2 if a.z <= 2 else 8
                rad = base - a.n_outer - total + a.charge
                if rad < 0:
                    rad = 0
                    if len(vals) > 1:
                        for v in vals:
                            if v - total >= 0:
                                rad = v - total
                                break
                rad2 = a.n_outer - total - a.charge
                if rad2 >= 0:
                    rad = min(rad, rad2)
                a.radicals = max(rad, 0)
            else:
                a.radicals = 0                     # metals / ions with an unconstrained valence list


def _more_electronegative(a, b):
    """RDKit's ordering: more outer electrons, then the lighter atom."""
    return a.n_outer > b.n_outer or (a.n_outer == b.n_outer and a.z < b.z)


def _pi_electrons(mol, i):
    """RDKit ``countAtomElec``: electrons an atom can put into a pi system, -1 when it cannot take part."""
    a = mol.atoms[i]
    vals = [v for v in a.valences if v >= 0]
    dv = vals[0] if vals else -1
    if dv <= 1:
        return -1
    degree = a.degree + a.total_h
    if degree > 3:
        return -1
    nlp = max(a.n_outer - dv - a.charge, 0)
    res = (dv - degree) + nlp - a.radicals
    if res > 1 and sum(mol.bonds[k].order for k in a.bonds) - a.degree > 1:
        res = 1
    return res


def _smallest_rings(mol):
    """For every ring bond the shortest cycle through it (breadth-first search avoiding the bond): the small-ring
    set Hueckel perception runs on (equals the SSSR for the ring systems of drug-like molecules)."""
    rings = {}
    for k, bond in enumerate(mol.bonds):
        if not bond.in_ring:
            continue
        prev = {bond.a: None}
        frontier = [bond.a]
        found = False
        while frontier and not found:
            nxt = []
            for v in frontier:
                for kk in mol.atoms[v].bonds:
                    if kk == k or not mol.bonds[kk].in_ring:
                        continue
                    w = mol.bonds[kk].other(v)
                    if w in prev:
                        continue
                    prev[w] = v
                    if w == bond.b:
                        found = True
                        break
                    nxt.append(w)
                if found:
                    break
            frontier = nxt
        if not found:
            continue
        path, v = [], bond.b
        while v is not None:
            path.append(v)
            v = prev[v]
        rings.setdefault(frozenset(path), path)
    return list(rings.values())


def _ring_bonds(mol, ring):
    out = []
    for x, y in zip(ring, ring[1:] + ring[:1]):
        for k in mol.atoms[x].bonds:
            if mol.bonds[k].other(x) == y:
                out.append(k)
                break
    return out


def _perceive_aromaticity(mol):
    """Hueckel perception for rings written in Kekulé form (RDKit's default model): every ring atom must be a donor
    — one electron from an atom with a ring double bond (or an exocyclic one to a less electronegative atom), none
    from an atom whose exocyclic double bond goes to a more electronegative one, two from a lone pair of an atom
    without multiple bonds — and a ring, or a fused set of rings (perimeter bonds), needs 4n+2 of them.  Atoms and
    bonds read as aromatic from lower-case SMILES keep their flag."""
    rings = _smallest_rings(mol)
    if not rings:
        return
    ring_atoms = set().union(*[set(r) for r in rings])
    donors = {}
    for i in ring_atoms:
        a = mol.atoms[i]
        ne = _pi_electrons(mol, i)
        if ne < 0:
            continue
        cyc = any(mol.bonds[k].order > 1 and mol.bonds[k].in_ring for k in a.bonds)
        exo = [k for k in a.bonds if mol.bonds[k].order > 1 and not mol.bonds[k].in_ring]
        multiple = cyc or bool(exo)
        if ne == 0:
            if exo or not multiple:
                donors[i] = 0
        elif ne == 1:
            if exo:
                other = mol.atoms[mol.bonds[exo[0]].other(i)]
                donors[i] = 0 if _more_electronegative(other, a) else 1
            elif multiple:
                donors[i] = 1
            elif a.charge == 1:
                donors[i] = 0
        else:
            if exo and _more_electronegative(mol.atoms[mol.bonds[exo[0]].other(i)], a):
                ne -= 1
            donors[i] = 1 if ne % 2 else 2
    cand = [r for r in rings if all(i in donors for i in r)]
    if not cand:
        return
    rbonds = [set(_ring_bonds(mol, r)) for r in cand]

    def huckel(atom_set):
        ne = sum(donors[i] for i in atom_set)
        return ne >= 2 and (ne - 2) % 4 == 0

    done_bonds = set()
    for r, bs in zip(cand, rbonds):
        if huckel(set(r)):
            for i in r:
                mol.atoms[i].aromatic = True
            for k in bs:
                mol.bonds[k].aromatic = True
            done_bonds |= bs
    # fused sets (azulene-like: no single ring passes, the perimeter does); bounded enumeration
    nr = len(cand)
    nbrs = [[j for j in range(nr) if j != i and rbonds[i] & rbonds[j]] for i in range(nr)]
    seen, budget = set(), 2000
    frontier = [frozenset([i]) for i in range(nr)]
    for _size in range(2, 7):
        nxt = []
        for sset in frontier:
            for i in sset:
                for j in nbrs[i]:
                    if j in sset:
                        continue
                    t = sset | {j}
                    if t in seen:
                        continue
                    seen.add(t)
                    budget -= 1
                    nxt.append(t)
                    count = {}
                    for q in t:
                        for k in rbonds[q]:
                            count[k] = count.get(k, 0) + 1
                    perimeter = {k for k, c in count.items() if c == 1}
                    if perimeter <= done_bonds:
                        continue
                    atoms_u = set().union(*[set(cand[q]) for q in t])
                    if huckel(atoms_u):
                        for i2 in atoms_u:
                            mol.atoms[i2].aromatic = True
                        for k in perimeter:
                            mol.bonds[k].aromatic = True
                        done_bonds |= perimeter
            if budget <= 0:
                break
        frontier = nxt
        if budget <= 0 or not frontier:
            break


def _conjugation(mol):
    """RDKit ``setConjugation``: aromatic bonds are conjugated; at an atom with 2-3 substituents, a bond of order
    >= 1.5 conjugates with every other bond whose far atom has at most 3 substituents and pi electrons to give."""
    for b in mol.bonds:
        b.conjugated = b.aromatic
    for i, a in enumerate(mol.atoms):
        sbo = a.degree + a.total_h
        if sbo < 2 or sbo > 3:
            continue
        for k1 in a.bonds:
            b1 = mol.bonds[k1]
            if not (b1.aromatic or b1.order >= 2):
                continue
            for k2 in a.bonds:
                if k2 == k1:
                    continue
                j = mol.bonds[k2].other(i)
                a2 = mol.atoms[j]
                if a2.degree + a2.total_h > 3:
                    continue
                if _pi_electrons(mol, j) > 0:
                    b1.conjugated = True
                    mol.bonds[k2].conjugated = True


def _hybridization(mol):
    """RDKit ``setHybridization``: orbitals = substituents + lone pairs (+ radicals below an octet); 4 orbitals on
    an atom with a conjugated bond and at most 3 neighbours count as SP2 (the O of an ester, the N of an amide)."""
    for a in mol.atoms:
        if a.z == 0:
            a.hybridization = OTHER
            continue
        deg = a.degree + a.total_h
        if a.z <= 1:
            norbs = deg
        else:
            total_valence = sum(mol.bonds[k].order for k in a.bonds) + a.total_h
            free = a.n_outer - (total_valence + a.charge)
            if total_valence + a.n_outer - a.charge < 8:
                norbs = deg + (free - a.radicals) // 2 + a.radicals
            else:
                norbs = deg + free // 2
        if norbs <= 1:
            a.hybridization = S
        elif norbs == 2:
            a.hybridization = SP
        elif norbs == 3:
            a.hybridization = SP2
        elif norbs == 4:
            conj = any(mol.bonds[k].conjugated for k in a.bonds)
            a.hybridization = SP3 if (a.degree > 3 or not conj) else SP2
        elif norbs == 5:
            a.hybridization = SP3D
        elif norbs == 6:
            a.hybridization = SP3D2
        else:
            a.hybridization = OTHER


def mol_from_smiles(smiles):
    """Parse + the sanitisation steps the 75 features depend on.  Raises :class:`SmilesError` where RDKit returns
    None (syntax, valence, Kekulé assignment)."""
    mol = parse_smiles(smiles)
    _fold_hydrogens(mol)
    _cleanup(mol)
    _mark_rings(mol)
    _kekulize(mol)
    _assign_hydrogens(mol)
    _perceive_aromaticity(mol)
    _conjugation(mol)
    _hybridization(mol)
    return mol


# ------------------------------------------------------------------------------------------------ features
def atom_features(mol, dtype=np.float64):
    """[n_atoms, 75] as ``graph_features.atom_features`` (bool_id_feat=False, explicit_H=False,
    use_chirality=False), float64 like the reference's ``np.array(results)`` of mixed bools / ints."""
    out = np.zeros((len(mol.atoms), N_ATOM_FEATURES), dtype=dtype)
    for i, a in enumerate(mol.atoms):
        out[i, _SYMBOL_INDEX.get(a.symbol, 43)] = 1
        if a.degree > 10:
            raise ValueError("input {0} not in allowable set{1}:".format(a.degree, list(range(11))))   # graph_features.py:34-36
        out[i, 44 + a.degree] = 1
        out[i, 55 + (a.implicit_h if a.implicit_h <= 6 else 6)] = 1
        out[i, 62] = a.charge
        out[i, 63] = a.radicals
        out[i, 64 + {SP: 0, SP2: 1, SP3: 2, SP3D: 3, SP3D2: 4}.get(a.hybridization, 4)] = 1
        out[i, 69] = 1 if a.aromatic else 0
        out[i, 70 + (a.total_h if a.total_h <= 4 else 4)] = 1
    return out


class ConvMolFeaturizer(object):
    """``deepchem.feat.ConvMolFeaturizer`` (graph_features.py:698-929) over the RDKit-free reader.  ``featurize``
    takes SMILES strings (or :class:`Mol`) and returns an object array of ``ConvMol``; a datapoint that fails
    yields an empty array in its slot, as in ``MolecularFeaturizer.featurize`` (base_classes.py:280-326)."""
    name = ['conv_mol']

    def __init__(self, master_atom=False, use_chirality=False, atom_properties=[], per_atom_fragmentation=False):
        if use_chirality or list(atom_properties):
            raise NotImplementedError("use_chirality / atom_properties need RDKit (CIP codes, property tables)")
        self.dtype = object
        self.master_atom, self.use_chirality = master_atom, use_chirality
        self.atom_properties, self.per_atom_fragmentation = list(atom_properties), per_atom_fragmentation

    def feature_length(self):
        return N_ATOM_FEATURES + len(self.atom_properties)

    def __hash__(self):
        return hash((self.master_atom, self.use_chirality, tuple(self.atom_properties)))

    def __eq__(self, other):
        return isinstance(self, other.__class__) and self.master_atom == other.master_atom and \
            self.use_chirality == other.use_chirality and tuple(self.atom_properties) == tuple(other.atom_properties)

    def _featurize(self, mol):
        nodes = atom_features(mol)
        adj = mol.adjacency_list()
        if self.master_atom:                                                 # graph_features.py:891-908
            nodes = np.concatenate([nodes, np.expand_dims(np.mean(nodes, axis=0), axis=0)], axis=0)
            adj.append([])
            fake = len(nodes) - 1
            for index in range(fake):
                adj[index].append(fake)
        if not self.per_atom_fragmentation:
            return ConvMol(nodes, adj)
        frags = []
        for i in range(nodes.shape[0]):                                       # graph_features.py:851-877
            new_a = [[v if v < i else v - 1 for v in pair if v != i] for j, pair in enumerate(adj) if j != i]
            frags.append(ConvMol(np.delete(nodes, i, axis=0), new_a))
        return frags

    def featurize(self, datapoints, log_every_n=1000, **kwargs):
        if isinstance(datapoints, (str, Mol)):
            datapoints = [datapoints]
        features = []
        for point in list(datapoints):
            try:
                mol = mol_from_smiles(point) if isinstance(point, str) else point
                features.append(self._featurize(mol))
            except Exception:
                features.append(np.array([]))
        out = np.empty(len(features), dtype=object)
        for i, f in enumerate(features):
            out[i] = f
        return out

    __call__ = featurize


def _featurize_chunk(smiles):
    """(features, adjacency) per readable string + the positions of the unreadable ones (worker of the pool below)."""
    mols, bad = [], []
    for i, s in enumerate(smiles):
        try:
            m = mol_from_smiles(s)
            mols.append((atom_features(m, np.float32), m.adjacency_list()))
        except Exception:
            bad.append(i)
    return mols, bad


def featurize_smiles_packed(smiles, n_jobs=1):
    """SMILES list -> ``PackedMols`` (the shard format the layout builder consumes), plus the indices of the
    strings that failed to parse (left out of the shard, as ``DataLoader`` drops failed datapoints).
    ``n_jobs`` > 1 reads contiguous chunks in that many worker processes (spawned, so a trainer that already holds a
    CUDA context is never forked — the calling script therefore needs the usual ``if __name__ == "__main__"`` guard);
    the result is identical to the serial one.  Starting the pool costs ~1 s, the chunks then scale with the workers
    (the reference's Tox21 file: 3.0 s serial, 0.95 s on 4 workers)."""
    from .synthetic import PackedMols
    smiles = list(smiles)
    n_jobs = max(1, min(int(n_jobs), (len(smiles) + 255) // 256))
    if n_jobs == 1:
        mols, bad = _featurize_chunk(smiles)
        return PackedMols.from_list(mols, n_feat=N_ATOM_FEATURES), bad
    import multiprocessing
    from concurrent.futures import ProcessPoolExecutor
    per = (len(smiles) + 4 * n_jobs - 1) // (4 * n_jobs)               # a few chunks per worker: even finish times
    starts = list(range(0, len(smiles), per))
    shards, bad = [], []
    with ProcessPoolExecutor(max_workers=n_jobs, mp_context=multiprocessing.get_context("spawn")) as pool:
        for lo, (arrays, b) in zip(starts, pool.map(_featurize_chunk_packed, [smiles[lo:lo + per] for lo in starts])):
            shards.append(PackedMols(*arrays))
            bad.extend(lo + i for i in b)
    return PackedMols.concat(shards), bad


def _featurize_chunk_packed(smiles):
    """One chunk as the four arrays of a PackedMols (a few large arrays cross the process boundary, not one small
    pair per molecule)."""
    from .synthetic import PackedMols
    mols, bad = _featurize_chunk(smiles)
    pm = PackedMols.from_list(mols, n_feat=N_ATOM_FEATURES)
    return (pm.atom_ptr, pm.adj_ptr, pm.adj_idx, pm.features), bad
