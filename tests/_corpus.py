"""Small deterministic text corpora for the CPU-side tests (python random, seed fixed)."""
import random

VOCAB = ("the of and to in that is was he for it with as his on be at by had not are but from or have an they "
         "which one you were her all she there would their we him been has when who will more no if out so said "
         "what up its about into than them can only other new some could time these two may then do first any my "
         "now such like our over man me even most made after also did many before must through back years where "
         "much your way well down should because each just those people how too little state good very make world "
         "still own see men work long get here between both life being under never day same another know while "
         "last might us great old year off come since against go came right used take three government "
         "governmental homogeneous approximate matching pattern string algorithm").split()


def make_text(nlines, seed=12345, paragraphs=False, trailing_newline=True, caps=0.1):
    rnd = random.Random(seed)
    out = []
    gap = rnd.randint(3, 8)
    for i in range(nlines):
        words = [rnd.choice(VOCAB) for _ in range(rnd.randint(6, 14))]
        line = " ".join(words)
        if rnd.random() < caps:
            line = line.capitalize()
        out.append(line)
        if paragraphs:
            gap -= 1
            if gap == 0:
                out.append("")
                if rnd.random() < 0.2:
                    out.append("")
                gap = rnd.randint(3, 8)
    s = "\n".join(out)
    if trailing_newline:
        s += "\n"
    return s.encode("ascii")


def mutate(rnd, s, nedits):
    s = list(s)
    for _ in range(nedits):
        op = rnd.randint(0, 2)
        p = rnd.randrange(len(s))
        if op == 0:
            s[p] = rnd.choice("abcdefghijklmnopqrstuvwxyz")
        elif op == 1 and len(s) > 2:
            del s[p]
        else:
            s.insert(p, rnd.choice("abcdefghijklmnopqrstuvwxyz"))
    return "".join(s)


def overlap_text(delim, seed=3):
    """records separated by a delimiter that overlaps itself, with chains of overlapping occurrences between some of them
    ("ababa", "abababa", "abaaba" for "aba"): no digits or colons, so that -n prefixes can be found in the output"""
    import random
    rnd = random.Random(seed)
    words = [w for w in make_text(300, seed=seed).decode().split() if w.isalpha()]
    d = delim
    per = next(p for p in range(1, len(d) + 1) if d[p:] == d[:len(d) - p])       # the delimiter's period
    out = []
    for i in range(400):
        out.append(" ".join(rnd.choice(words) for _ in range(rnd.randint(0, 6))))
        r = rnd.random()
        chain = d if r < 0.5 else d + d[len(d) - per:] * rnd.randint(1, 4) if r < 0.8 else d + d if r < 0.9 else d[:per] + d
        out.append(chain)
    return "".join(out).encode()
