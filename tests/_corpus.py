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
