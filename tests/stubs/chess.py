"""Test stand-in for the `chess` (python-chess) package, which is not installed in this image: just enough of its
API for the reference's callers (game_session.py, game_state.py, lichess_session.py, calibrate_sensitivity.py) to run
in the unchanged-caller smoke tests.  Moves are pseudo-legal (no check detection, castling or en passant) -- exact
for the opening plies the tests play.  Test infrastructure only."""
WHITE, BLACK = True, False
PAWN, KNIGHT, BISHOP, ROOK, QUEEN, KING = 1, 2, 3, 4, 5, 6
SQUARES = list(range(64))
_SYM = {PAWN: "p", KNIGHT: "n", BISHOP: "b", ROOK: "r", QUEEN: "q", KING: "k"}


def square(file_index, rank_index):
    return rank_index * 8 + file_index


def square_file(sq):
    return sq & 7


def square_rank(sq):
    return sq >> 3


def square_name(sq):
    return "abcdefgh"[sq & 7] + str((sq >> 3) + 1)


class Piece:
    def __init__(self, piece_type, color):
        self.piece_type, self.color = piece_type, color

    def symbol(self):
        s = _SYM[self.piece_type]
        return s.upper() if self.color else s


class Move:
    def __init__(self, from_square, to_square, promotion=None):
        self.from_square, self.to_square, self.promotion = from_square, to_square, promotion

    def uci(self):
        return square_name(self.from_square) + square_name(self.to_square) + (_SYM[self.promotion] if self.promotion else "")

    @classmethod
    def from_uci(cls, uci):
        f = square("abcdefgh".index(uci[0]), int(uci[1]) - 1)
        t = square("abcdefgh".index(uci[2]), int(uci[3]) - 1)
        promo = {v: k for k, v in _SYM.items()}[uci[4]] if len(uci) > 4 else None
        return cls(f, t, promo)

    def __eq__(self, o):
        return isinstance(o, Move) and (self.from_square, self.to_square, self.promotion) == (o.from_square, o.to_square, o.promotion)

    def __hash__(self):
        return hash((self.from_square, self.to_square, self.promotion))

    def __repr__(self):
        return "Move.from_uci(%r)" % self.uci()

    def __bool__(self):
        return True


class _LegalMoves:
    def __init__(self, board):
        self.b = board

    def __iter__(self):
        return iter(self.b._moves())

    def __contains__(self, m):
        return m in self.b._moves()

    def __len__(self):
        return len(self.b._moves())


class Board:
    def __init__(self):
        self.reset()

    def reset(self):
        self._p = {}
        back = [ROOK, KNIGHT, BISHOP, QUEEN, KING, BISHOP, KNIGHT, ROOK]
        for f in range(8):
            self._p[square(f, 0)] = Piece(back[f], WHITE); self._p[square(f, 1)] = Piece(PAWN, WHITE)
            self._p[square(f, 7)] = Piece(back[f], BLACK); self._p[square(f, 6)] = Piece(PAWN, BLACK)
        self.turn = WHITE
        self.move_stack = []

    def __bool__(self):
        return True

    def piece_at(self, sq):
        return self._p.get(sq)

    def piece_map(self):
        return dict(self._p)

    @property
    def legal_moves(self):
        return _LegalMoves(self)

    def _moves(self):
        out = []
        for sq, p in self._p.items():
            if p.color != self.turn:
                continue
            f, r = sq & 7, sq >> 3
            if p.piece_type == PAWN:
                d, start, last = (1, 1, 7) if p.color else (-1, 6, 0)
                def add(to):
                    if (to >> 3) == last:
                        out.extend(Move(sq, to, pr) for pr in (QUEEN, ROOK, BISHOP, KNIGHT))
                    else:
                        out.append(Move(sq, to))
                if 0 <= r + d < 8 and square(f, r + d) not in self._p:
                    add(square(f, r + d))
                    if r == start and square(f, r + 2 * d) not in self._p:
                        out.append(Move(sq, square(f, r + 2 * d)))
                for df in (-1, 1):
                    if 0 <= f + df < 8 and 0 <= r + d < 8:
                        t = self._p.get(square(f + df, r + d))
                        if t is not None and t.color != p.color:
                            add(square(f + df, r + d))
                continue
            if p.piece_type in (KNIGHT, KING):
                steps = ([(1, 2), (2, 1), (2, -1), (1, -2), (-1, -2), (-2, -1), (-2, 1), (-1, 2)] if p.piece_type == KNIGHT
                         else [(1, 0), (1, 1), (0, 1), (-1, 1), (-1, 0), (-1, -1), (0, -1), (1, -1)])
                for df, dr in steps:
                    if 0 <= f + df < 8 and 0 <= r + dr < 8:
                        t = self._p.get(square(f + df, r + dr))
                        if t is None or t.color != p.color:
                            out.append(Move(sq, square(f + df, r + dr)))
                continue
            rays = []
            if p.piece_type in (BISHOP, QUEEN):
                rays += [(1, 1), (-1, 1), (1, -1), (-1, -1)]
            if p.piece_type in (ROOK, QUEEN):
                rays += [(1, 0), (-1, 0), (0, 1), (0, -1)]
            for df, dr in rays:
                x, y = f + df, r + dr
                while 0 <= x < 8 and 0 <= y < 8:
                    t = self._p.get(square(x, y))
                    if t is None:
                        out.append(Move(sq, square(x, y)))
                    else:
                        if t.color != p.color:
                            out.append(Move(sq, square(x, y)))
                        break
                    x += df; y += dr
        return out

    def is_capture(self, move):
        return move.to_square in self._p

    def is_en_passant(self, move):
        return False

    def push(self, move):
        p = self._p.pop(move.from_square)
        if move.promotion:
            p = Piece(move.promotion, p.color)
        self._p[move.to_square] = p
        self.move_stack.append(move)
        self.turn = not self.turn

    def push_uci(self, uci):
        m = Move.from_uci(uci)
        if m not in self.legal_moves:
            raise ValueError("illegal uci: " + uci)
        self.push(m)
        return m

    def peek(self):
        return self.move_stack[-1]

    def fen(self):
        rows = []
        for r in range(7, -1, -1):
            row, gap = "", 0
            for f in range(8):
                p = self._p.get(square(f, r))
                if p is None:
                    gap += 1
                else:
                    row += (str(gap) if gap else "") + p.symbol(); gap = 0
            rows.append(row + (str(gap) if gap else ""))
        return "/".join(rows) + (" w" if self.turn else " b") + " - - 0 %d" % (len(self.move_stack) // 2 + 1)

    def set_fen(self, fen):
        self._p = {}
        rows = fen.split()[0].split("/")
        inv = {v: k for k, v in _SYM.items()}
        for i, row in enumerate(rows):
            f = 0
            for ch in row:
                if ch.isdigit():
                    f += int(ch)
                else:
                    self._p[square(f, 7 - i)] = Piece(inv[ch.lower()], ch.isupper()); f += 1
        parts = fen.split()
        self.turn = len(parts) < 2 or parts[1] == "w"
        self.move_stack = []
