"""CPU tests of the host-side mirror of the reference's operator interface (prophylo_b200/algorithm.py): the parts of
lib/PhyloProf/Profile.pm and lib/PhyloProf/Algorithm/ppp.pm that run before the GPU call."""
import pytest

from prophylo_b200.algorithm import Profile, collapse, perl_num, perl_split_tab, perl_true, ppp


class MockTaxonomy:
    """stands in for SeqToolBox::Taxonomy: collapse_taxon(id, rank) from a table (Taxonomy.pm:296-360)"""

    def __init__(self, table):
        self.table = table

    def collapse_taxon(self, taxon, rank):
        return self.table.get(rank, {}).get(taxon) or None


def test_profile_file_parse_mirrors_profile_pm(tmp_path):
    f = tmp_path / "q.profile"
    f.write_text("# comment\n10\t1\n11\t0\n10\t0\n11\t1\n0\t1\n12\t0\t\n")
    q = Profile(file=str(f))
    # a non-zero value is never overwritten, a zero one is (Profile.pm:463-468); taxon "0" is Perl-false and dropped (:460);
    # a trailing empty field is dropped by split (:452)
    assert q.get_hash() == {"10": "1", "11": "1", "12": "0"}
    assert (q.get_yes(), q.get_no(), q.get_total()) == (2, 1, 3)
    # -header skips PHYSICAL line 1 whatever it holds (Profile.pm:447-449)
    g = tmp_path / "h.profile"
    g.write_text("Taxon\tValue\textra\n# c\n5\t1\n")
    assert Profile(file=str(g), header=True).get_hash() == {"5": "1"}
    with pytest.raises(ValueError, match="must contain only two fields"):
        Profile(file=str(g))  # without -header line 1 has three fields
    e = tmp_path / "e.profile"
    e.write_text("5\t1\n\n6\t0\n")
    with pytest.raises(ValueError, match="must contain only two fields"):
        Profile(file=str(e))  # an empty line splits into zero fields: the reference confesses (:454-456)
    with pytest.raises(OSError, match="Can' open"):
        Profile(file=str(tmp_path / "missing"))


def test_perl_scalar_helpers():
    assert [perl_num(x) for x in ("1", "0", "abc", "1e3x", " .5", "-2.5", 3)] == [1.0, 0.0, 0.0, 1000.0, 0.5, -2.5, 3.0]
    assert [perl_true(x) for x in ("0", "", None, "00", "0.0", "a")] == [False, False, False, True, True, True]
    assert perl_split_tab("a\t1\t") == ["a", "1"] and perl_split_tab("") == [] and perl_split_tab("a\t\tb") == ["a", "", "b"]


def test_rank_collapse_and_croak():
    tax = MockTaxonomy({"family": {"s1": "f1"}, "genus": {"s2": "g2"}})
    assert collapse(tax, "family", "s1") == "f1"  # the rank itself
    assert collapse(tax, "family", "s2") == "g2"  # fallback to genus (ppp.pm:354-358)
    assert collapse(tax, "family", "s3") == "s3"  # fallback to the raw id (:360-362)
    assert collapse(None, None, "s1") == "s1"     # no -rank: identity, no taxonomy needed
    a = ppp(Profile(profile={"f1": 1, "x": 0}), Profile(raw=["s1", "s3"], yes=2, no=0), rank="family")
    with pytest.raises(RuntimeError, match="Taxonomy does not exist"):  # ppp.pm:345-349, before any GPU work
        a.get_match_score()
    # the query profile collapses its own taxa when given -rank / -taxonomy (Profile.pm:455-466, 516-549)
    q = Profile(profile={"s1": 1, "s2": 0, "s3": 0}, rank="family", taxonomy=tax)
    assert q.get_hash() == {"f1": 1, "g2": 0, "s3": 0}


def test_constructor_errors_mirror_ppp_pm():
    with pytest.raises(ValueError, match="Parameter is undefined"):
        ppp(None, None)
    with pytest.raises(TypeError, match="Parameter mismatch"):
        ppp({"a": 1}, ["a"])
    a = ppp(Profile(profile={}), Profile(raw=["a"], yes=1, no=0))
    with pytest.raises(ZeroDivisionError):
        a.get_match_score()  # total == 0 and no -prob: "Illegal division by zero" (ppp.pm:141)
