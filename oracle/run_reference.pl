#!/usr/bin/perl
# oracle/run_reference.pl -- TEST INFRASTRUCTURE ONLY.
# Drives the UNMODIFIED reference scorer (lib/PhyloProf/Algorithm/ppp.pm:130-245, loaded from
# $REFERENCE_ROOT/lib, default /root/reference/lib) on JSON cases and prints one JSON result per
# case with %.17g doubles.  Needs oracle/stubs on @INC and the XS shim (see oracle/Makefile `ref`).
# Used by oracle/gen_golden.py to pin the C restatement and to generate tests/golden/*.json;
# never runs on the GPU box (the reference tree does not exist there).
#
# input (stdin): JSON array of cases
#   { "query": {taxon: 0|1, ...}, "target": [taxon, ...] | {taxon: rank, ...},
#     "prob": optional number, "dup": optional 0|1, "file": optional path of a profile file for the query,
#     "rank": optional rank name (-rank), "collapse": optional {rank: {taxon: collapsed taxon}} served by a mock
#     -taxonomy object exposing collapse_taxon (what ppp.pm:328-370 calls); "rank" without "collapse" passes no taxonomy }
# output (file argv[0], else stdout): JSON array of [score, yes, number, depth, p] as strings ("%.17g") / ints
use strict;
use warnings;
use JSON::PP;
use FindBin;
BEGIN {
    my $root = $ENV{REFERENCE_ROOT} || "/root/reference";
    unshift @INC, "$FindBin::Bin/stubs", "$root/lib";
    $ENV{CEPHES_SHIM_SO} ||= "$FindBin::Bin/_ref/Cephes_shim.so";
    $ENV{CEPHES_REAL_SO} ||= "$FindBin::Bin/_ref/Cephes.so";
}
use PhyloProf::Profile;
use PhyloProf::Algorithm::ppp;

{
    package MockTaxonomy;    # stands in for SeqToolBox::Taxonomy: collapse_taxon(id, rank) from a table (Taxonomy.pm:296-360)
    sub new { my ( $c, $map ) = @_; return bless { map => $map }, $c }
    sub collapse_taxon { my ( $s, $id, $rank ) = @_; my $v = $s->{map}{$rank}{$id}; return $v ? $v : undef }
}

my $cases = do { local $/; decode_json(<STDIN>) };
my @out;
open( my $saved, ">&", \*STDOUT ) or die;
foreach my $c (@$cases) {
    my $prof1;
    if ( $c->{file} ) {
        $prof1 = PhyloProf::Profile->new( -file => $c->{file} );
    } else {
        $prof1 = PhyloProf::Profile->new( -profile => $c->{query} );
    }
    my $t = $c->{target};
    my $n = ref($t) eq 'ARRAY' ? scalar(@$t) : scalar( keys %$t );
    my $prof2 = PhyloProf::Profile->new( -raw => $t, -yes => $n, -no => 0 );
    my @args = ( -prof1 => $prof1, -prof2 => $prof2 );
    push @args, ( -prob => $c->{prob} ) if defined $c->{prob};
    push @args, ( -dup  => 1 )          if $c->{dup};
    push @args, ( -rank => $c->{rank} ) if $c->{rank};
    push @args, ( -taxonomy => MockTaxonomy->new( $c->{collapse} ) ) if $c->{collapse};
    my $alg = PhyloProf::Algorithm::ppp->new(@args);
    # Cephes' mtherr prints "bdtrc domain error" on stdout: keep it out of the JSON stream
    open( STDOUT, ">", "/dev/null" ) or die;
    my ( $score, $yes, $number, $depth, $p ) = eval { no warnings; $alg->get_match_score() };
    my $err = $@;
    open( STDOUT, ">&", $saved ) or die;
    if ($err) { $err =~ s/\s+$//; push @out, { error => $err }; next; }
    push @out,
        [ sprintf( "%.17g", $score ), $yes + 0, $number + 0, $depth + 0, sprintf( "%.17g", $p ),
          $prof1->get_yes() + 0, $prof1->get_total() + 0 ];
}
# Cephes' printf output is flushed by C stdio at exit on fd 1; the JSON goes to the file named by argv[0]
my $dest = shift @ARGV;
if ($dest) { open( my $fh, ">", $dest ) or die "$dest: $!"; print $fh encode_json( \@out ), "\n"; close $fh; }
else       { print encode_json( \@out ), "\n"; }
