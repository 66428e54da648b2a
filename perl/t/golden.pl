#!/usr/bin/perl
# perl/t/golden.pl -- GPU test of the Perl drop-in: PhyloProf::Algorithm::ppp_b200->new(...)->get_match_score() on the
# committed golden cases (tests/golden/ppp_list_cases.json, produced by the reference itself): ARRAY targets with
# duplicates and foreign taxa, HASH targets, -prob, -dup.  yes/number/depth exact, score within 1e-9 on -log p.
use strict;
use warnings;
use FindBin;
use lib "$FindBin::Bin/../lib", "$FindBin::Bin/lib";
use JSON::PP;
use PhyloProf::Profile;
use PhyloProf::Algorithm::ppp_b200;

my $file = shift || "$FindBin::Bin/../../tests/golden/ppp_list_cases.json";
open( my $fh, "<", $file ) or die "$file: $!";
my $g = decode_json( do { local $/; <$fh> } );
my ( $n, $bad, $skipped ) = ( 0, 0, 0 );
foreach my $c ( @{ $g->{known_answers} }, @{ $g->{fuzz} } ) {
    my $t = $c->{target};
    my $cnt = ref($t) eq 'ARRAY' ? scalar(@$t) : scalar( keys %$t );
    my $p1 = PhyloProf::Profile->new( -profile => $c->{query} );
    my $p2 = PhyloProf::Profile->new( -raw => $t, -yes => $cnt, -no => 0 );
    my @args = ( -prof1 => $p1, -prof2 => $p2 );
    push @args, ( -prob => $c->{prob} ) if defined $c->{prob};
    if ( $c->{dup} ) { push @args, ( -dup => 1 ); $skipped++; }    # counted below as -dup cases run
    my @r = PhyloProf::Algorithm::ppp_b200->new(@args)->get_match_score();
    my @e = @{ $c->{expect} };
    $n++;
    my $ok = $r[1] == $e[1] && $r[2] == $e[2] && $r[3] == $e[3] && $r[4] == $e[4];
    if ($ok) {
        my ( $a, $b ) = ( $r[0], $e[0] + 0 );
        if ( $a != $b ) {
            $ok = 0 if $a <= 0 || $b <= 0;
            $ok = 0 if $ok && abs( log($a) - log($b) ) > 1e-9 * abs( log($b) );
        }
    }
    unless ($ok) { $bad++; print "MISMATCH $c->{name}: got @r expected @e\n" if $bad < 10; }
}
# batch entry with ranked lists: the KA-2 query against three target lists in one call
{
    my $ka = $g->{known_answers}[1];
    my @universe = sort keys %{ $ka->{query} };
    my $gpu = PhyloProf::B200->new( -device => $ENV{PPP_B200_DEVICE} || 0 );
    my $hits = $gpu->score_ranked( \@universe, [ $ka->{query} ], [ $ka->{target}, [ reverse @{ $ka->{target} } ], [] ] );
    my @e = @{ $ka->{expect} };
    my $h = $hits->[0];
    unless ( @$hits == 3 && $h->[3] == $e[1] && $h->[4] == $e[2] && $h->[5] == $e[3] && $hits->[2][2] == 1 ) {
        $bad++;
        print "MISMATCH ranked batch: @$h vs @e\n";
    }
}
# threshold run + device-side cut-offs through the XS binding: rows arrive sorted, CUTOFF as ppp_cutoff.pl computes it
{
    my $ka = $g->{known_answers}[1];
    my @universe = sort keys %{ $ka->{query} };
    my $gpu = PhyloProf::B200->new( -device => $ENV{PPP_B200_DEVICE} || 0 );
    my @t = @{ $ka->{target} };
    my @lists = map { [ @t[ 0 .. $_ ] ] } ( 2 .. $#t );
    my $hits = $gpu->score_ranked( \@universe, [ $ka->{query} ], \@lists, -filter => 'thresh', -thresh => -0.0005 );
    my $cuts = $gpu->printed_cutoffs( 1, 20 );
    my ( $first, $rows, $marker, $cutoff, $status ) = @{ $cuts->[0] };
    my @logs = map { sprintf( "%.3f", log( $_->[2] ) / log(10) ) } @$hits;
    my ( $last, $want );
    foreach my $s ( sort { $b <=> $a } grep { $_ <= -0.001 } @logs ) {    # bin/ppp_cutoff.pl:130-148
        next if $s > -1;
        unless ($last) { $last = $s; next; }
        if ( abs( ( $last - $s ) / $last * 100 ) > 20 ) { $want = $last; last; }
        $last = $s;
    }
    my $sorted = !grep { $hits->[$_][2] < $hits->[ $_ - 1 ][2] } ( 1 .. $#$hits );
    my $same = defined($want) ? ( defined($cutoff) && $cutoff == sprintf( "%.0f", $want * 1000 ) ) : !defined($cutoff);
    unless ( $rows == @$hits && $first == 0 && $sorted && ( $status != 0 || $same ) ) {
        $bad++;
        print "MISMATCH printed_cutoffs: rows $rows vs ", scalar(@$hits), " cutoff ", ( $cutoff // 'undef' ), " want ", ( $want // 'undef' ), "\n";
    }
}
# -rank (ppp.pm:328-370): the reference run with a mock -taxonomy object (oracle/gen_golden_rank.py); croak without one
{
    package MockTaxonomy;
    sub new { my ( $c, $map ) = @_; return bless { map => $map }, $c }
    sub collapse_taxon { my ( $s, $id, $rank ) = @_; my $v = $s->{map}{$rank}{$id}; return $v ? $v : undef }
}
my $nrank = 0;
{
    ( my $rf = $file ) =~ s/ppp_list_cases\.json$/ppp_rank_cases.json/;
    open( my $rh, "<", $rf ) or die "$rf: $!";
    my $rg = decode_json( do { local $/; <$rh> } );
    foreach my $c ( @{ $rg->{cases} } ) {
        my $t = $c->{target};
        my $cnt = ref($t) eq 'ARRAY' ? scalar(@$t) : scalar( keys %$t );
        my @args = ( -prof1 => PhyloProf::Profile->new( -profile => $c->{query} ), -prof2 => PhyloProf::Profile->new( -raw => $t, -yes => $cnt, -no => 0 ),
                     -rank => $c->{rank}, -taxonomy => MockTaxonomy->new( $c->{collapse} ) );
        push @args, ( -prob => $c->{prob} ) if defined $c->{prob};
        push @args, ( -dup => 1 ) if $c->{dup};
        my @r = PhyloProf::Algorithm::ppp_b200->new(@args)->get_match_score();
        my @e = @{ $c->{expect} };
        $nrank++;
        my $ok = $r[1] == $e[1] && $r[2] == $e[2] && $r[3] == $e[3] && $r[4] == $e[4];
        if ( $ok && $r[0] != $e[0] + 0 ) {
            $ok = 0 if $r[0] <= 0 || $e[0] <= 0 || abs( log( $r[0] ) - log( $e[0] ) ) > 1e-9 * abs( log( $e[0] ) );
        }
        unless ($ok) { $bad++; print "MISMATCH $c->{name}: got @r expected @e\n" if $bad < 10; }
    }
    my $c = $rg->{croak};
    my $alg = PhyloProf::Algorithm::ppp_b200->new( -prof1 => PhyloProf::Profile->new( -profile => $c->{query} ),
                                                   -prof2 => PhyloProf::Profile->new( -raw => $c->{target}, -yes => 2, -no => 0 ), -rank => $c->{rank} );
    eval { $alg->get_match_score() };
    unless ( $@ =~ /\Q$c->{error}\E/ ) { $bad++; print "MISMATCH: -rank without -taxonomy did not croak ($@)\n"; }
    # callers also poke the taxonomy in after construction (bin/ppp.pl:572)
    $alg->{_dbh} = MockTaxonomy->new( {} );
    my @r = eval { $alg->get_match_score() };
    if ($@) { $bad++; print "MISMATCH: poked _dbh not honoured ($@)\n"; }
}
# p comes from the WHOLE query profile (ppp.pm:138-142) even when some of its taxa lie outside the universe, and a false
# -prob counts as absent (ppp.pm:100-102)
{
    my $gpu = PhyloProf::B200->new( -device => $ENV{PPP_B200_DEVICE} || 0 );
    my %q = ( a => 1, b => 1, c => 0, d => 0, zz1 => 0, zz2 => 0, zz3 => 1, zz4 => 0 );    # 3 yes of 8
    my @universe = qw(a b c d);
    my $ref = $gpu->score_ranked( \@universe, [ \%q ], [ [qw(a b c)] ], -probs => [ 3 / 8 ] );
    my $got = $gpu->score_ranked( \@universe, [ \%q ], [ [qw(a b c)] ] );
    my $got0 = $gpu->score_ranked( \@universe, [ \%q ], [ [qw(a b c)] ], -prob => 0 );
    my $bits = $gpu->score_bits( \@universe, [ \%q ], [ [qw(a b c)] ] );
    foreach my $h ( $got, $got0, $bits ) {
        unless ( $h->[0][2] == $ref->[0][2] && $h->[0][3] == 2 ) { $bad++; print "MISMATCH p of a profile with taxa outside the universe: $h->[0][2] vs $ref->[0][2]\n"; }
    }
}
print "perl drop-in: $n cases ($skipped with -dup), $nrank -rank cases, $bad mismatches\n";
exit( $bad ? 1 : 0 );
