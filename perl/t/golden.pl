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
print "perl drop-in: $n cases ($skipped with -dup), $bad mismatches\n";
exit( $bad ? 1 : 0 );
